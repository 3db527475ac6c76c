// Internal interface between translation units of libpsgb200.so: psg_mixct.cu holds the compile-time mixed-radix
// kernels for round FFT lengths (sti_mixct.cuh, plan list mixct_plans.inc).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include "sti_common.cuh"

struct MixctInfo {
    int threads;   // per CTA
    int groups;    // frames a CTA works on at once
    size_t smem;   // dynamic shared memory per CTA
    int occ;       // resident CTAs per SM
    char name[48];
};
// a compile-time plan exists for this length / sample type; fills `info` for the form that suits frames_per_col
// (one frame group per CTA for single-frame columns, the plan's default otherwise).  Returns a cudaError_t
// (cudaErrorInvalidValue: no plan).
int psg_mixct_query(int n, int iq_type, int frames_per_col, MixctInfo* info);
// enqueue: `a` complete (chunk / nsplit / partial set)
int psg_mixct_launch(int n, int iq_type, int frames_per_col, const StiArgs& a, long long grid, cudaStream_t st);
