// Internal interface between the two translation units of libpsgb200.so: psg_r32.cu holds the radix-32
// whole-frame kernels (sti_r32.cuh), psg_b200.cu everything else (separate units keep the edit-compile loop of
// one kernel family short and build in parallel).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include "sti_common.cuh"

// resident clusters (CTAs for nfft = 16384) of the kernel for this size / sample type on `device`;
// 0 when the device cannot co-schedule the cluster.  Returns a cudaError_t.
int psg_r32_max_groups(int logn, int iq_type, int device, int sms, int* ngroups);
// enqueue the kernel: `a` complete (chunk / nsplit / partial set), nitems = ncol * nsub * nsplit
int psg_r32_launch(int logn, int iq_type, const StiArgs& a, int nitems, int ngroups, cudaStream_t st);
// threads per CTA and cluster size of the kernel psg_r32_launch would run for this size
void psg_r32_describe(int logn, int* threads, int* cl);
