// libpsgb200.so -- host side of the C ABI declared in include/psg_b200.h.
//
// Owns: plan tables (window, twiddles), kernel-variant dispatch and launch geometry, the
// time-median launch, and the host-buffer entry point.  No torch, no cuFFT, no CPU fallback: if
// there is no sm_100 device every compute entry point fails with PSG_ERR_NODEVICE.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/psg_b200.h"
#include "sti_kernels.cuh"
#include "sti_cluster.cuh"
#include "sti_whole.cuh"
#include "sti_whole16.cuh"
#include "sti_bluestein.cuh"
#include "psg_r32.h"
#include "psg_mixct.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
// Debugging / tuning knobs (psg_debug_set_*): they apply to the CALLING THREAD only -- the viewer runs up to seven
// worker threads over one library (drfview.py:177-178), and one of them forcing a variant must not change the
// kernel another one is about to launch.  (The wrappers keep the std::atomic / mutex interface of the code below.)
template <class T>
struct TlsKnob {
    T v;
    T load() const { return v; }
    void store(T x) { v = x; }
};
struct NoMutex {
    void lock() {}
    void unlock() {}
};
static thread_local TlsKnob<int> g_force_generic{0};
static thread_local NoMutex g_variant_mu;
static thread_local std::string g_variant_override;  // "" = automatic
static thread_local TlsKnob<int> g_items_per_slot{24};  // work items per resident CTA slot the column split aims for
static thread_local TlsKnob<int> g_use_multi{1};        // Mode R: use the multi-column twin kernels
// Scratch cap of the large-nfft split path.  Measured (profiles/r01_sweep_split_scratch*.txt): L2-sized
// chunks (16..128 MiB) lose more to the three short dependent launches per chunk than they save in
// HBM traffic; 1..4 GiB chunks run each phase at its own roofline.
static thread_local TlsKnob<long long> g_split_scratch_bytes{2048ll << 20};
// psg_sti_host: recordings whose touched span exceeds this are streamed in column chunks of about this size
static thread_local TlsKnob<long long> g_host_chunk_bytes{1024ll << 20};
static thread_local TlsKnob<int> g_cluster_rowtma{0};  // cluster path: 0 = rows loaded to registers (default), 1 = by bulk copy, 2 = DSMEM exchange

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(PSG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                   \
    } while (0)

// Every entry point works on the plan's device and leaves the caller's current device as it found it (torch
// derives ITS current device from cudaGetDevice: a call for cuda:1 from a thread working on cuda:0 must not
// redirect that thread's later allocations).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) ok = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define PSG_ON_DEVICE(dev)                                                                   \
    DeviceGuard dev_guard_(dev);                                                             \
    if (!dev_guard_.ok) return fail(PSG_ERR_CUDA, "cudaSetDevice(%d) failed (%s:%d)", (dev), __FILE__, __LINE__)

// ------------------------------------------------------------------------------------------------
// kernel variants
// ------------------------------------------------------------------------------------------------
struct Variant {
    const char* name;
    int logn, F, loader, threads, minb, iqt, twp, multi;
    size_t smem;
    const void* fn;
};

template <int LOGN, int E, int R0, int R1, int R2, int R3, int F, int LOADER, int STAGES, int XBUF, int MINB, int PFL2 = 0,
          int IQT = IQ_C64, int TWP = 0, int MULTI = 0>
static Variant make_variant(const char* name) {
    using CF = FusedCfg<LOGN, E, R0, R1, R2, R3, F, LOADER, STAGES, XBUF, IQT, TWP>;
    Variant v;
    v.name = name;
    v.logn = LOGN;
    v.F = F;
    v.loader = LOADER;
    v.threads = CF::NT;
    v.minb = MINB;
    v.iqt = IQT;
    v.twp = TWP;
    v.multi = MULTI;
    v.smem = CF::smem_bytes;
    v.fn = (const void*)sti_fused_kernel<LOGN, E, R0, R1, R2, R3, F, LOADER, STAGES, XBUF, MINB, PFL2, IQT, TWP, MULTI>;
    return v;
}

#define L PSG_LOADER_LDG
#define M PSG_LOADER_TMA
// name = <loader><logn>_<radices>_f<F>_s<stages>x<xbuf>; first match per (logn, loader) is default
static const Variant g_variants[] = {
    //            LOGN E  R0  R1  R2 R3  F  LD ST XB MINB
    make_variant<5, 8, 4, 8, 1, 1, 32, L, 1, 2, 8>("ldg5_4x8_f32"),
    make_variant<6, 8, 8, 8, 1, 1, 32, L, 1, 2, 4>("ldg6_8x8_f32"),
    make_variant<7, 16, 8, 16, 1, 1, 16, L, 1, 2, 4>("ldg7_8x16_f16"),
    make_variant<8, 16, 16, 16, 1, 1, 8, L, 1, 2, 4>("ldg8_16x16_f8"),
    make_variant<9, 8, 8, 8, 8, 1, 1, L, 1, 2, 16>("ldg9_8x8x8_f1"),
    make_variant<10, 16, 4, 16, 16, 1, 1, L, 1, 2, 8>("ldg10_4x16x16_f1"),
    make_variant<10, 16, 4, 16, 16, 1, 1, M, 2, 1, 8>("tma10_4x16x16_f1_s2x1"),
    make_variant<11, 16, 8, 16, 16, 1, 1, L, 1, 2, 4>("ldg11_8x16x16_f1"),
    make_variant<11, 16, 8, 16, 16, 1, 1, M, 2, 1, 4>("tma11_8x16x16_f1_s2x1"),
    make_variant<12, 16, 16, 16, 16, 1, 1, M, 2, 1, 2>("tma12_16x16x16_f1_s2x1"),
    make_variant<12, 16, 16, 16, 16, 1, 1, M, 1, 2, 2>("tma12_16x16x16_f1_s1x2"),
    make_variant<12, 16, 16, 16, 16, 1, 1, L, 1, 2, 2>("ldg12_16x16x16_f1"),
    make_variant<12, 16, 16, 16, 16, 1, 1, L, 1, 1, 2, 2>("ldg12_16x16x16_f1_x1_pf2"),
    make_variant<12, 16, 16, 16, 16, 1, 1, M, 2, 1, 2, 0, IQ_C64, 1>("tma12_16x16x16_f1_s2x1_tp"),
    make_variant<12, 16, 16, 16, 16, 1, 1, M, 2, 1, 2, 0, IQ_C64, 2>("tma12_16x16x16_f1_s2x1_tq"),
    make_variant<12, 16, 16, 16, 16, 1, 1, M, 1, 2, 2, 0, IQ_C64, 2>("tma12_16x16x16_f1_s1x2_tq"),
    make_variant<11, 16, 8, 16, 16, 1, 1, M, 1, 2, 4, 0, IQ_C64, 2>("tma11_8x16x16_f1_s1x2_tq"),
    make_variant<10, 16, 8, 8, 16, 1, 1, M, 2, 1, 8, 0, IQ_C64, 2>("tma10_8x8x16_f1_s2x1_tq"),
    make_variant<10, 16, 4, 16, 16, 1, 1, M, 1, 2, 8, 0, IQ_C64, 2>("tma10_4x16x16_f1_s1x2_tq"),
    make_variant<9, 8, 8, 8, 8, 1, 1, M, 2, 1, 16, 0, IQ_C64, 1>("tma9_8x8x8_f1_s2x1_tp"),
    make_variant<9, 16, 2, 16, 16, 1, 1, M, 2, 1, 16, 0, IQ_C64, 2>("tma9_2x16x16_f1_s2x1_tq"),
    make_variant<8, 16, 16, 16, 1, 1, 2, M, 2, 1, 16>("tma8_16x16_f2_s2x1"),
    make_variant<8, 16, 16, 16, 1, 1, 4, M, 2, 1, 8>("tma8_16x16_f4_s2x1"),
    make_variant<9, 16, 2, 16, 16, 1, 1, M, 1, 2, 16, 0, IQ_C64, 2>("tma9_2x16x16_f1_s1x2_tq"),
    // 512 points by one warp: radix-2 pass in registers on the end of pass 0, one shared-memory round trip
    make_variant<9, 16, 16, 2, 16, 1, 1, M, 1, 2, 16, 0, IQ_C64, 2>("tma9_16x2x16_f1_s1x2_tq"),
    make_variant<9, 16, 16, 2, 16, 1, 1, M, 2, 1, 16, 0, IQ_C64, 2>("tma9_16x2x16_f1_s2x1_tq"),
    make_variant<7, 16, 8, 16, 1, 1, 4, M, 2, 1, 16>("tma7_8x16_f4_s2x1"),
    make_variant<7, 16, 8, 16, 1, 1, 8, M, 2, 1, 8>("tma7_8x16_f8_s2x1"),
    make_variant<6, 8, 8, 8, 1, 1, 8, M, 2, 1, 16>("tma6_8x8_f8_s2x1"),
    make_variant<6, 8, 8, 8, 1, 1, 16, M, 2, 1, 8>("tma6_8x8_f16_s2x1"),
    make_variant<5, 8, 4, 8, 1, 1, 16, M, 2, 1, 16>("tma5_4x8_f16_s2x1"),
    make_variant<5, 8, 4, 8, 1, 1, 32, M, 2, 1, 8>("tma5_4x8_f32_s2x1"),
    make_variant<11, 16, 8, 16, 16, 1, 1, M, 2, 1, 4, 0, IQ_C64, 2>("tma11_8x16x16_f1_s2x1_tq"),
    make_variant<10, 16, 4, 16, 16, 1, 1, M, 2, 1, 8, 0, IQ_C64, 2>("tma10_4x16x16_f1_s2x1_tq"),
    make_variant<13, 16, 16, 8, 8, 8, 1, M, 2, 1, 1, 0, IQ_C64, 2>("tma13_16x8x8x8_f1_s2x1_tq"),
    // radix-2 pass in registers on the end of pass 1 (lane pairs exchange by SHFL): three shared-memory passes
    make_variant<13, 16, 16, 16, 2, 16, 1, M, 2, 1, 1, 0, IQ_C64, 2>("tma13_16x16x2x16_f1_s2x1_tq"),
    make_variant<13, 16, 16, 16, 2, 16, 1, M, 1, 2, 1, 0, IQ_C64, 2>("tma13_16x16x2x16_f1_s1x2_tq"),
    make_variant<11, 16, 8, 16, 16, 1, 1, M, 2, 1, 4, 0, IQ_C64, 1>("tma11_8x16x16_f1_s2x1_tp"),
    make_variant<10, 16, 4, 16, 16, 1, 1, M, 2, 1, 8, 0, IQ_C64, 1>("tma10_4x16x16_f1_s2x1_tp"),
    make_variant<13, 16, 16, 8, 8, 8, 1, M, 2, 1, 1, 0, IQ_C64, 1>("tma13_16x8x8x8_f1_s2x1_tp"),
    make_variant<9, 8, 8, 8, 8, 1, 1, L, 1, 2, 16, 0, IQ_C64, 1>("ldg9_8x8x8_f1_tp"),
    make_variant<10, 16, 4, 16, 16, 1, 1, L, 1, 2, 8, 0, IQ_C64, 1>("ldg10_4x16x16_f1_tp"),
    make_variant<11, 16, 8, 16, 16, 1, 1, L, 1, 2, 4, 0, IQ_C64, 1>("ldg11_8x16x16_f1_tp"),
    make_variant<12, 16, 16, 16, 16, 1, 1, L, 1, 2, 2, 0, IQ_C64, 1>("ldg12_16x16x16_f1_tp"),
    make_variant<12, 16, 16, 16, 16, 1, 1, L, 1, 1, 2, 2, IQ_C64, 1>("ldg12_16x16x16_f1_x1_pf2_tp"),
    make_variant<13, 16, 2, 16, 16, 16, 1, L, 1, 1, 1, 0, IQ_C64, 1>("ldg13_2x16x16x16_f1_tp"),
    make_variant<12, 8, 8, 8, 8, 8, 1, M, 2, 1, 2>("tma12_8x8x8x8_f1_s2x1"),
    make_variant<13, 16, 16, 8, 8, 8, 1, M, 2, 1, 1>("tma13_16x8x8x8_f1_s2x1"),
    make_variant<13, 16, 8, 8, 8, 16, 1, M, 2, 1, 1>("tma13_8x8x8x16_f1_s2x1"),
    make_variant<13, 16, 2, 16, 16, 16, 1, L, 1, 1, 1>("ldg13_2x16x16x16_f1"),
    // one-frame-per-column twins (_m) of the default TMA variants: several column blocks per CTA
    make_variant<5, 8, 4, 8, 1, 1, 32, M, 2, 1, 8, 0, IQ_C64, 0, 1>("tma5_4x8_f32_s2x1_m"),
    make_variant<6, 8, 8, 8, 1, 1, 16, M, 2, 1, 8, 0, IQ_C64, 0, 1>("tma6_8x8_f16_s2x1_m"),
    make_variant<7, 16, 8, 16, 1, 1, 8, M, 2, 1, 8, 0, IQ_C64, 0, 1>("tma7_8x16_f8_s2x1_m"),
    make_variant<8, 16, 16, 16, 1, 1, 4, M, 2, 1, 8, 0, IQ_C64, 0, 1>("tma8_16x16_f4_s2x1_m"),
    make_variant<9, 16, 2, 16, 16, 1, 1, M, 1, 2, 16, 0, IQ_C64, 2, 1>("tma9_2x16x16_f1_s1x2_tq_m"),
    make_variant<9, 16, 16, 2, 16, 1, 1, M, 1, 2, 16, 0, IQ_C64, 2, 1>("tma9_16x2x16_f1_s1x2_tq_m"),
    make_variant<9, 16, 16, 2, 16, 1, 1, M, 2, 1, 16, 0, IQ_C64, 2, 1>("tma9_16x2x16_f1_s2x1_tq_m"),
    make_variant<10, 16, 4, 16, 16, 1, 1, M, 1, 2, 8, 0, IQ_C64, 2, 1>("tma10_4x16x16_f1_s1x2_tq_m"),
    make_variant<11, 16, 8, 16, 16, 1, 1, M, 1, 2, 4, 0, IQ_C64, 2, 1>("tma11_8x16x16_f1_s1x2_tq_m"),
    make_variant<12, 16, 16, 16, 16, 1, 1, M, 2, 1, 2, 0, IQ_C64, 2, 1>("tma12_16x16x16_f1_s2x1_tq_m"),
    make_variant<13, 16, 16, 8, 8, 8, 1, M, 2, 1, 1, 0, IQ_C64, 2, 1>("tma13_16x8x8x8_f1_s2x1_tq_m"),
    make_variant<13, 16, 16, 16, 2, 16, 1, M, 2, 1, 1, 0, IQ_C64, 2, 1>("tma13_16x16x2x16_f1_s2x1_tq_m"),
    // raw integer IQ ingest (complex int16 / int8): the default geometry of every size, both loaders
    make_variant<5, 8, 4, 8, 1, 1, 32, L, 1, 2, 8, 0, IQ_CI16>("ldg5_4x8_f32_i16"),
    make_variant<6, 8, 8, 8, 1, 1, 32, L, 1, 2, 4, 0, IQ_CI16>("ldg6_8x8_f32_i16"),
    make_variant<7, 16, 8, 16, 1, 1, 16, L, 1, 2, 4, 0, IQ_CI16>("ldg7_8x16_f16_i16"),
    make_variant<8, 16, 16, 16, 1, 1, 8, L, 1, 2, 4, 0, IQ_CI16>("ldg8_16x16_f8_i16"),
    make_variant<9, 8, 8, 8, 8, 1, 1, L, 1, 2, 16, 0, IQ_CI16, 1>("ldg9_8x8x8_f1_tp_i16"),
    make_variant<10, 16, 4, 16, 16, 1, 1, M, 2, 1, 8, 0, IQ_CI16, 2>("tma10_4x16x16_f1_s2x1_tq_i16"),
    make_variant<10, 16, 4, 16, 16, 1, 1, L, 1, 2, 8, 0, IQ_CI16, 1>("ldg10_4x16x16_f1_tp_i16"),
    make_variant<11, 16, 8, 16, 16, 1, 1, M, 2, 1, 4, 0, IQ_CI16, 2>("tma11_8x16x16_f1_s2x1_tq_i16"),
    make_variant<11, 16, 8, 16, 16, 1, 1, L, 1, 2, 4, 0, IQ_CI16, 1>("ldg11_8x16x16_f1_tp_i16"),
    make_variant<12, 16, 16, 16, 16, 1, 1, M, 2, 1, 2, 0, IQ_CI16, 2>("tma12_16x16x16_f1_s2x1_tq_i16"),
    make_variant<12, 16, 16, 16, 16, 1, 1, L, 1, 2, 2, 0, IQ_CI16, 1>("ldg12_16x16x16_f1_tp_i16"),
    make_variant<13, 16, 16, 8, 8, 8, 1, M, 2, 1, 1, 0, IQ_CI16, 2>("tma13_16x8x8x8_f1_s2x1_tq_i16"),
    make_variant<13, 16, 16, 16, 2, 16, 1, M, 2, 1, 1, 0, IQ_CI16, 2>("tma13_16x16x2x16_f1_s2x1_tq_i16"),
    make_variant<13, 16, 2, 16, 16, 16, 1, L, 1, 1, 1, 0, IQ_CI16, 1>("ldg13_2x16x16x16_f1_tp_i16"),
    make_variant<5, 8, 4, 8, 1, 1, 32, L, 1, 2, 8, 0, IQ_CI8>("ldg5_4x8_f32_i8"),
    make_variant<6, 8, 8, 8, 1, 1, 32, L, 1, 2, 4, 0, IQ_CI8>("ldg6_8x8_f32_i8"),
    make_variant<7, 16, 8, 16, 1, 1, 16, L, 1, 2, 4, 0, IQ_CI8>("ldg7_8x16_f16_i8"),
    make_variant<8, 16, 16, 16, 1, 1, 8, L, 1, 2, 4, 0, IQ_CI8>("ldg8_16x16_f8_i8"),
    make_variant<9, 8, 8, 8, 8, 1, 1, L, 1, 2, 16, 0, IQ_CI8, 1>("ldg9_8x8x8_f1_tp_i8"),
    make_variant<10, 16, 4, 16, 16, 1, 1, M, 2, 1, 8, 0, IQ_CI8, 2>("tma10_4x16x16_f1_s2x1_tq_i8"),
    make_variant<10, 16, 4, 16, 16, 1, 1, L, 1, 2, 8, 0, IQ_CI8, 1>("ldg10_4x16x16_f1_tp_i8"),
    make_variant<11, 16, 8, 16, 16, 1, 1, M, 2, 1, 4, 0, IQ_CI8, 2>("tma11_8x16x16_f1_s2x1_tq_i8"),
    make_variant<11, 16, 8, 16, 16, 1, 1, L, 1, 2, 4, 0, IQ_CI8, 1>("ldg11_8x16x16_f1_tp_i8"),
    make_variant<12, 16, 16, 16, 16, 1, 1, M, 2, 1, 2, 0, IQ_CI8, 2>("tma12_16x16x16_f1_s2x1_tq_i8"),
    make_variant<12, 16, 16, 16, 16, 1, 1, L, 1, 2, 2, 0, IQ_CI8, 1>("ldg12_16x16x16_f1_tp_i8"),
    make_variant<13, 16, 16, 8, 8, 8, 1, M, 2, 1, 1, 0, IQ_CI8, 2>("tma13_16x8x8x8_f1_s2x1_tq_i8"),
    make_variant<13, 16, 16, 16, 2, 16, 1, M, 2, 1, 1, 0, IQ_CI8, 2>("tma13_16x16x2x16_f1_s2x1_tq_i8"),
    make_variant<13, 16, 2, 16, 16, 16, 1, L, 1, 1, 1, 0, IQ_CI8, 1>("ldg13_2x16x16x16_f1_tp_i8"),
};
#undef L
#undef M
static const int g_nvariants = (int)(sizeof(g_variants) / sizeof(g_variants[0]));

static const Variant* variant_by_name(const char* name) {
    for (int i = 0; i < g_nvariants; ++i)
        if (strcmp(name, g_variants[i].name) == 0) return &g_variants[i];
    return nullptr;
}

// Default variant per FFT length, from the measured sweeps (profiles/r01_sweep_*.txt, 4 and 12 GB of
// IQ, full-coverage Mode A): the TMA ring wins at every size once the CTAs are small (one frame group
// from 512 up, 4..32 frames per CTA below); mid-pass twiddles rebuilt from register-resident
// W^1,2,4,8 (_tq) or one row load (_tp) instead of 15 loads win everywhere (the kernels are LSU-bound,
// not HBM- or FMA-bound).
static const char* const g_default_tma[] = {"tma5_4x8_f32_s2x1", "tma6_8x8_f16_s2x1", "tma7_8x16_f8_s2x1", "tma8_16x16_f4_s2x1",
                                            "tma9_16x2x16_f1_s2x1_tq", "tma10_4x16x16_f1_s1x2_tq", "tma11_8x16x16_f1_s1x2_tq",
                                            "tma12_16x16x16_f1_s2x1_tq", "tma13_16x16x2x16_f1_s2x1_tq"};
static const char* const g_default_ldg[] = {"ldg5_4x8_f32", "ldg6_8x8_f32", "ldg7_8x16_f16", "ldg8_16x16_f8", "ldg9_8x8x8_f1_tp",
                                            "ldg10_4x16x16_f1_tp", "ldg11_8x16x16_f1_tp", "ldg12_16x16x16_f1_tp",
                                            "ldg13_2x16x16x16_f1_tp"};
// raw integer IQ (suffix _i16 / _i8 appended): the instantiated subset
static const char* const g_default_tma_int[] = {"ldg5_4x8_f32", "ldg6_8x8_f32", "ldg7_8x16_f16", "ldg8_16x16_f8", "ldg9_8x8x8_f1_tp",
                                                "tma10_4x16x16_f1_s2x1_tq", "tma11_8x16x16_f1_s2x1_tq",
                                                "tma12_16x16x16_f1_s2x1_tq", "tma13_16x16x2x16_f1_s2x1_tq"};

static const Variant* pick_variant(int logn, bool tma_ok, int iqt) {
    {
        std::lock_guard<NoMutex> lk(g_variant_mu);
        if (!g_variant_override.empty()) {
            const Variant* v = variant_by_name(g_variant_override.c_str());
            if (v && v->logn == logn && v->iqt == iqt && (v->loader != PSG_LOADER_TMA || tma_ok)) return v;
        }
    }
    if (logn < 5 || logn > 13) return nullptr;
    std::string name = (tma_ok ? (iqt == IQ_C64 ? g_default_tma : g_default_tma_int) : g_default_ldg)[logn - 5];
    if (iqt == IQ_CI16) name += "_i16";
    if (iqt == IQ_CI8) name += "_i8";
    return variant_by_name(name.c_str());
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct psg_plan {
    int nfft, logn, device, sms;
    float* d_win = nullptr;
    float2* d_tw = nullptr;
    float2* d_twp = nullptr;
    std::vector<float> h_win;
    std::vector<double> h_win_d;  // w/sum(w) in float64 (Bluestein folds it into a complex table)
    // large-nfft split path (nfft = r0 * 4096): first-pass twiddles, the 4096-point sub-transform's
    // tables, the L2-sized scratch of first-pass outputs, the sub-spectra and carried sums
    float2* d_twa = nullptr;
    float* d_ones = nullptr;
    float2* d_twp_sub = nullptr;
    float2* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    float* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    float* d_carry = nullptr;
    size_t carry_bytes = 0;
    // cluster path: resident clusters per kernel instantiation (0 = not queried, -1 = unavailable)
    float2* d_xslot = nullptr;
    size_t xslot_bytes = 0;
    // arbitrary nfft (Bluestein): convolution length 2^logm, tables
    int logm = 0;
    float2* d_aw = nullptr;
    float2* d_bbr = nullptr;
    float2* d_twm = nullptr;
    // high-radix Bluestein (sti_bluestein.cuh, M <= 16384): B in the position order of the radix plan, full W_M table
    float2* d_bpos = nullptr;
    float2* d_twf = nullptr;
    int bs_npass = 0, bs_radix[4] = {0, 0, 0, 0};
    // direct mixed-radix transform for nfft = 2^a 3^b 5^c (sti_mixed_kernel): radices, 0 passes = not applicable
    int mx_npass = 0, mx_radix[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long* d_colb = nullptr;
    size_t colb_bytes = 0;
    std::vector<long long> h_colb;
    long long colb_stride = 0;
    // scratch (grow-only; a plan is used by one host thread at a time)
    float* d_partial = nullptr;
    size_t partial_elems = 0;
    float2* d_gwork = nullptr;
    float* d_gacc = nullptr;
    size_t gwork_slabs = 0;
    // psg_sti_host staging
    void* d_in = nullptr;
    size_t in_bytes = 0;
    void* d_in2 = nullptr;  // second staging buffer, copy stream and events of the streamed host path
    size_t in2_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    long long* d_off = nullptr;
    size_t off_elems = 0;
    long long* d_offchk = nullptr;  // psg_sti_run_checked: the column offsets clamped to the addressable range
    size_t offchk_elems = 0;
    float* d_out[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t out_elems[4] = {0, 0, 0, 0};
    cudaStream_t stream = nullptr;
    std::vector<char> attr_done;  // per variant: smem attribute set on this device
    std::vector<int> occ;         // per variant: resident CTAs per SM
    char variant_name[64];
    std::string twp_name;  // variant whose pass tables d_twp holds
};

static double bessel_i0(double x) {
    // power series sum ((x/2)^(2k) / (k!)^2); converges to double precision for |x| < ~700
    const double q = 0.25 * x * x;
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 500; ++k) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < sum * 1e-17) break;
    }
    return sum;
}

static int ilog2_exact(int n) {
    int l = 0;
    while ((1 << l) < n) ++l;
    return ((1 << l) == n) ? l : -1;
}

static int check_device(int device, int* sms) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(PSG_ERR_NODEVICE, "no CUDA device: %s", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(PSG_ERR_ARG, "device %d out of range (have %d)", device, n);
    cudaDeviceProp pr;
    CUDA_TRY(cudaGetDeviceProperties(&pr, device));
    if (pr.major != 10)
        return fail(PSG_ERR_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    pr.major, pr.minor);
    *sms = pr.multiProcessorCount;
    return PSG_OK;
}

extern "C" int psg_version(void) { return PSG_ABI_VERSION; }
extern "C" const char* psg_last_error(void) { return g_err; }
extern "C" int psg_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(PSG_ERR_NODEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}
extern "C" int64_t psg_launch_count(void) { return g_launches.load(); }
extern "C" int psg_debug_set_force_generic(int on) {
    g_force_generic.store(on ? 1 : 0);
    return PSG_OK;
}
extern "C" int psg_debug_set_variant(const char* name) {
    std::lock_guard<NoMutex> lk(g_variant_mu);
    g_variant_override = name ? name : "";
    if (!g_variant_override.empty()) {
        bool ok = g_variant_override == "split" || g_variant_override == "cluster" || g_variant_override == "cluster_ldg" ||
                  g_variant_override == "cluster_dsmem" || g_variant_override == "whole" || g_variant_override == "whole_s2" ||
                  g_variant_override == "whole_s8" || g_variant_override == "whole_f" || g_variant_override == "whole_r2" || g_variant_override == "whole_r4" ||
                  g_variant_override == "bluestein_r2" || g_variant_override == "bluestein" || g_variant_override == "r32" ||
                  g_variant_override == "mixed_rt";
        for (int i = 0; i < g_nvariants; ++i) ok = ok || g_variant_override == g_variants[i].name;
        if (!ok) {
            g_variant_override.clear();
            return fail(PSG_ERR_ARG, "unknown kernel variant '%s'", name);
        }
    }
    return PSG_OK;
}
extern "C" int psg_debug_set_split_scratch(int64_t bytes) {
    if (bytes < (1 << 20)) return fail(PSG_ERR_ARG, "psg_debug_set_split_scratch: at least 1 MiB");
    g_split_scratch_bytes.store(bytes);
    return PSG_OK;
}
extern "C" int psg_debug_set_host_chunk(int64_t bytes) {
    if (bytes < (1 << 16)) return fail(PSG_ERR_ARG, "psg_debug_set_host_chunk: at least 64 KiB");
    g_host_chunk_bytes.store(bytes);
    return PSG_OK;
}
extern "C" int psg_debug_set_mode_r_multi(int on) {
    g_use_multi.store(on ? 1 : 0);
    return PSG_OK;
}
extern "C" int psg_debug_set_items_per_slot(int n) {
    if (n < 1 || n > 1024) return fail(PSG_ERR_ARG, "psg_debug_set_items_per_slot: 1..1024");
    g_items_per_slot.store(n);
    return PSG_OK;
}
extern "C" int psg_variant_count(void) { return g_nvariants; }
extern "C" const char* psg_variant_name(int i) { return (i >= 0 && i < g_nvariants) ? g_variants[i].name : ""; }
extern "C" int psg_variant_logn(int i) { return (i >= 0 && i < g_nvariants) ? g_variants[i].logn : -1; }

// Host-side window table w[n]/sum(w) (fp64 math, fp32 result) without touching a device: what the
// plan uploads.  Exposed so CPU-only tests can pin the table against scipy's.
static int window_table_d(int nfft, int window_kind, double beta, std::vector<double>& w, double* sum_out) {
    if (nfft < 1) return fail(PSG_ERR_ARG, "psg_window_table: bad arguments");
    w.resize(nfft);
    if (window_kind == PSG_WINDOW_KAISER) {
        const double a = 0.5 * nfft, i0b = bessel_i0(beta);
        for (int n = 0; n < nfft; ++n) {
            const double r = (n - a) / a;
            const double arg = 1.0 - r * r;
            w[n] = (nfft == 1) ? 1.0 : bessel_i0(beta * sqrt(arg > 0 ? arg : 0.0)) / i0b;
        }
    } else if (window_kind == PSG_WINDOW_BOXCAR) {
        for (int n = 0; n < nfft; ++n) w[n] = 1.0;
    } else {
        return fail(PSG_ERR_ARG, "unknown window kind %d", window_kind);
    }
    double s = 0.0;  // pairwise-ish: numpy's sum is pairwise; plain Kahan keeps us within 1 ulp of it
    double c = 0.0;
    for (int n = 0; n < nfft; ++n) {
        const double y = w[n] - c, t = s + y;
        c = (t - s) - y;
        s = t;
    }
    for (int n = 0; n < nfft; ++n) w[n] /= s;
    if (sum_out) *sum_out = s;
    return PSG_OK;
}

extern "C" int psg_window_table(int nfft, int window_kind, double beta, float* host_out, double* sum_out) {
    if (!host_out) return fail(PSG_ERR_ARG, "psg_window_table: bad arguments");
    std::vector<double> w;
    int rc = window_table_d(nfft, window_kind, beta, w, sum_out);
    if (rc) return rc;
    for (int n = 0; n < nfft; ++n) host_out[n] = (float)w[n];
    return PSG_OK;
}

// in-place radix-2 FFT in float64 on the host (plan tables only)
static void host_fft(std::vector<double>& re, std::vector<double>& im) {
    const size_t n = re.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = -2.0 * M_PI / (double)len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const double wr = cos(ang * (double)k), wi = sin(ang * (double)k);
                const size_t a = i + k, b = a + len / 2;
                const double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] += xr; im[a] += xi;
            }
    }
}

// Bluestein tables for a non power-of-two nfft (see sti_kernels.cuh): aw, bit-reversed B/M, W_M
static int build_bluestein(psg_plan* p) {
    const int N = p->nfft;
    int logm = 1;
    while ((1ll << logm) < 2ll * N - 1) ++logm;
    const size_t M = (size_t)1 << logm;
    p->logm = logm;
    std::vector<double> cr(N), ci(N);
    for (int n = 0; n < N; ++n) {
        const long long q = ((long long)n * n) % (2ll * N);  // exact phase reduction
        const double ang = M_PI * (double)q / (double)N;
        cr[n] = cos(ang);
        ci[n] = sin(ang);
    }
    std::vector<float2> aw(N), bbr(M), twm(M / 2);
    for (int n = 0; n < N; ++n) aw[n] = make_float2((float)(p->h_win_d[n] * cr[n]), (float)(-p->h_win_d[n] * ci[n]));
    std::vector<double> br(M, 0.0), bi(M, 0.0);
    for (int n = 0; n < N; ++n) {
        br[n] = cr[n]; bi[n] = ci[n];
        if (n) { br[M - n] = cr[n]; bi[M - n] = ci[n]; }
    }
    host_fft(br, bi);
    for (size_t i = 0; i < M; ++i) {
        size_t r = 0;
        for (int b = 0; b < logm; ++b) r |= ((i >> b) & 1) << (logm - 1 - b);
        bbr[i] = make_float2((float)(br[r] / (double)M), (float)(bi[r] / (double)M));
    }
    for (size_t m = 0; m < M / 2; ++m) {
        const double ang = -2.0 * M_PI * (double)m / (double)M;
        twm[m] = make_float2((float)cos(ang), (float)sin(ang));
    }
    if (logm <= 14) {
        // radix plan: 2^(logm % 4) first, then 16s; after the forward passes position pos = sum_q k_q S_q holds
        // frequency k_0 + R_0 k_1 + R_0 R_1 k_2 + ... (index algebra of sti_kernels.cuh)
        int np = 0;
        if (logm % 4) p->bs_radix[np++] = 1 << (logm % 4);
        for (int i = 0; i < logm / 4; ++i) p->bs_radix[np++] = 16;
        p->bs_npass = np;
        std::vector<float2> bpos(M), twf(M);
        for (size_t pos = 0; pos < M; ++pos) {
            size_t rem = pos, s = M, freq = 0, mul = 1;
            for (int q = 0; q < np; ++q) {
                s /= (size_t)p->bs_radix[q];
                freq += (rem / s) * mul;
                rem %= s;
                mul *= (size_t)p->bs_radix[q];
            }
            bpos[pos] = make_float2((float)(br[freq] / (double)M), (float)(bi[freq] / (double)M));
        }
        for (size_t m = 0; m < M; ++m) {
            const double ang = -2.0 * M_PI * (double)m / (double)M;
            twf[m] = make_float2((float)cos(ang), (float)sin(ang));
        }
        CUDA_TRY(cudaMalloc(&p->d_bpos, sizeof(float2) * M));
        CUDA_TRY(cudaMalloc(&p->d_twf, sizeof(float2) * M));
        CUDA_TRY(cudaMemcpy(p->d_bpos, bpos.data(), sizeof(float2) * M, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->d_twf, twf.data(), sizeof(float2) * M, cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaMalloc(&p->d_aw, sizeof(float2) * N));
    CUDA_TRY(cudaMalloc(&p->d_bbr, sizeof(float2) * M));
    CUDA_TRY(cudaMalloc(&p->d_twm, sizeof(float2) * (M / 2)));
    CUDA_TRY(cudaMemcpy(p->d_aw, aw.data(), sizeof(float2) * N, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(p->d_bbr, bbr.data(), sizeof(float2) * M, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(p->d_twm, twm.data(), sizeof(float2) * (M / 2), cudaMemcpyHostToDevice));
    return PSG_OK;
}

extern "C" int psg_plan_create(psg_plan** out, int nfft, int window_kind, double beta, int device) {
    if (!out) return fail(PSG_ERR_ARG, "psg_plan_create: out is NULL");
    *out = nullptr;
    if (nfft < PSG_MIN_NFFT || nfft > PSG_MAX_NFFT)
        return fail(PSG_ERR_UNSUPPORTED, "nfft=%d outside [%d, %d]", nfft, PSG_MIN_NFFT, PSG_MAX_NFFT);
    const int logn = ilog2_exact(nfft);  // -1: not a power of two -> Bluestein
    int sms = 0;
    int rc = check_device(device, &sms);
    if (rc) return rc;
    PSG_ON_DEVICE(device);

    psg_plan* p = new psg_plan();
    p->nfft = nfft;
    p->logn = logn;
    p->device = device;
    p->sms = sms;
    p->h_win.resize(nfft);
    p->attr_done.assign(g_nvariants, 0);
    p->occ.assign(g_nvariants, 0);
    rc = window_table_d(nfft, window_kind, beta, p->h_win_d, nullptr);
    if (rc) { delete p; return rc; }
    for (int n = 0; n < nfft; ++n) p->h_win[n] = (float)p->h_win_d[n];

    // full twiddle table (generic kernels) and the per-pass tables of every tuned variant layout
    std::vector<float2> tw(nfft);
    for (int m = 0; m < nfft; ++m) {
        const double ang = -2.0 * M_PI * (double)m / (double)nfft;
        tw[m] = make_float2((float)cos(ang), (float)sin(ang));
    }
    cudaError_t e;
#define PLAN_TRY(expr)                                                                             \
    if ((e = (expr)) != cudaSuccess) {                                                             \
        psg_plan_destroy(p);                                                                       \
        return fail(PSG_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e));                  \
    }
    PLAN_TRY(cudaMalloc(&p->d_win, sizeof(float) * nfft));
    PLAN_TRY(cudaMalloc(&p->d_tw, sizeof(float2) * nfft));
    PLAN_TRY(cudaMemcpy(p->d_win, p->h_win.data(), sizeof(float) * nfft, cudaMemcpyHostToDevice));
    PLAN_TRY(cudaMemcpy(p->d_tw, tw.data(), sizeof(float2) * nfft, cudaMemcpyHostToDevice));
    PLAN_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
#undef PLAN_TRY
    if (logn < 0) {
        rc = build_bluestein(p);
        if (rc) { psg_plan_destroy(p); return rc; }
        // 2^a 3^b 5^c 7^d 11^e 13^f that fits shared memory: radix plan of the direct transform (power-of-two part
        // first, then the primes from the largest down)
        int rem = nfft, a2 = 0, n3 = 0, n5 = 0, n7 = 0, n11 = 0, n13 = 0;
        while (rem % 2 == 0) { rem /= 2; ++a2; }
        while (rem % 3 == 0) { rem /= 3; ++n3; }
        while (rem % 5 == 0) { rem /= 5; ++n5; }
        while (rem % 7 == 0) { rem /= 7; ++n7; }
        while (rem % 11 == 0) { rem /= 11; ++n11; }
        while (rem % 13 == 0) { rem /= 13; ++n13; }
        const int np = (a2 % 4 ? 1 : 0) + a2 / 4 + n3 + n5 + n7 + n11 + n13;
        if (rem == 1 && np <= 10 && (size_t)(psg_pad(nfft) + 4) * 8 + (size_t)(nfft + 4) * 4 <= 200 * 1024) {
            int k = 0;
            if (a2 % 4) p->mx_radix[k++] = 1 << (a2 % 4);
            for (int i = 0; i < a2 / 4; ++i) p->mx_radix[k++] = 16;
            for (int i = 0; i < n13; ++i) p->mx_radix[k++] = 13;
            for (int i = 0; i < n11; ++i) p->mx_radix[k++] = 11;
            for (int i = 0; i < n7; ++i) p->mx_radix[k++] = 7;
            for (int i = 0; i < n5; ++i) p->mx_radix[k++] = 5;
            for (int i = 0; i < n3; ++i) p->mx_radix[k++] = 3;
            p->mx_npass = k;
        }
    }
    p->variant_name[0] = 0;
    *out = p;
    return PSG_OK;
}

extern "C" int psg_plan_destroy(psg_plan* p) {
    if (!p) return PSG_OK;
    DeviceGuard dev_guard_(p->device);
    cudaFree(p->d_win);
    cudaFree(p->d_tw);
    cudaFree(p->d_twp);
    cudaFree(p->d_twa);
    cudaFree(p->d_ones);
    cudaFree(p->d_twp_sub);
    cudaFree(p->d_scratch);
    cudaFree(p->d_tmp);
    cudaFree(p->d_carry);
    cudaFree(p->d_xslot);
    cudaFree(p->d_colb);
    cudaFree(p->d_aw);
    cudaFree(p->d_bbr);
    cudaFree(p->d_twm);
    cudaFree(p->d_bpos);
    cudaFree(p->d_twf);
    cudaFree(p->d_partial);
    cudaFree(p->d_gwork);
    cudaFree(p->d_gacc);
    cudaFree(p->d_in);
    cudaFree(p->d_in2);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    for (int i = 0; i < 2; ++i) {
        if (p->ev_copied[i]) cudaEventDestroy(p->ev_copied[i]);
        if (p->ev_free[i]) cudaEventDestroy(p->ev_free[i]);
    }
    cudaFree(p->d_off);
    cudaFree(p->d_offchk);
    for (int i = 0; i < 4; ++i) cudaFree(p->d_out[i]);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return PSG_OK;
}

extern "C" int psg_plan_nfft(const psg_plan* p) { return p ? p->nfft : fail(PSG_ERR_ARG, "plan is NULL"); }

extern "C" int psg_plan_window(const psg_plan* p, float* host_out) {
    if (!p || !host_out) return fail(PSG_ERR_ARG, "psg_plan_window: NULL argument");
    memcpy(host_out, p->h_win.data(), sizeof(float) * p->nfft);
    return PSG_OK;
}

// per-pass twiddle tables for radices (r[0..np)) -- layouts documented in sti_kernels.cuh:
// pass 0 and mid passes with stride > 32 use the column layout, mid passes with stride <= 32 the
// row layout (R+2 complex per row, entry k of row n' = W^{n'*k}).
static std::vector<float2> build_pass_tables(int n, const int* r, int np, int twp) {
    std::vector<float2> t;
    int s = n;
    for (int p = 0; p + 1 < np; ++p) {
        s /= r[p];
        const int m = r[p] * s;
        auto w = [&](int i, int k) {
            const double ang = -2.0 * M_PI * (double)((long long)i * k % m) / (double)m;
            return make_float2((float)cos(ang), (float)sin(ang));
        };
        const bool row = p >= 1 && s <= 32;
        if (p >= 1 && twp) {
            // power layout: only W^1, W^2, W^4, W^8 (exponents below the radix) are stored
            const int npw = psg_npow(r[p]);
            if (row) {
                for (int i = 0; i < s; ++i)
                    for (int q = 0; q < 6; ++q) t.push_back(q < npw ? w(i, 1 << q) : make_float2(0.f, 0.f));
            } else {
                for (int q = 0; q < npw; ++q)
                    for (int i = 0; i < s; ++i) t.push_back(w(i, 1 << q));
            }
        } else if (row) {
            for (int i = 0; i < s; ++i)
                for (int k = 0; k < r[p] + 2; ++k) t.push_back(k < r[p] ? w(i, k) : make_float2(0.f, 0.f));
        } else {
            for (int k = 1; k < r[p]; ++k)
                for (int i = 0; i < s; ++i) t.push_back(w(i, k));
        }
    }
    if (t.empty()) t.push_back(make_float2(1.f, 0.f));
    return t;
}

// radices of a variant, parsed from its name ("..._16x8x8_...")
static int variant_radices(const Variant* v, int* r) {
    const char* s = strchr(v->name, '_');
    int np = 0;
    if (!s) return 0;
    ++s;
    while (*s && np < 4) {
        r[np++] = atoi(s);
        while (*s >= '0' && *s <= '9') ++s;
        if (*s != 'x') break;
        ++s;
    }
    return np;
}

static int ensure_buffer(void** ptr, size_t* have, size_t want_bytes) {
    if (*have >= want_bytes && *ptr) return PSG_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *have = 0;
    cudaError_t e = cudaMalloc(ptr, want_bytes);
    if (e != cudaSuccess) return fail(PSG_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", want_bytes, cudaGetErrorString(e));
    *have = want_bytes;
    return PSG_OK;
}

static int floor_pow2(int x) {
    int p = 1;
    while (p * 2 <= x) p *= 2;
    return p;
}

extern "C" const char* psg_plan_variant(const psg_plan* p) { return p ? p->variant_name : ""; }

// Launch one tuned fused kernel (+ the fixed-order finalize when columns are split over CTAs).
// `a` carries everything but the launch geometry; min_iters = fewest frame iterations worth a
// work item of its own (bounds the share of the per-item epilogue).
static int variant_ready(psg_plan* p, const Variant* v) {
    const int vi = (int)(v - g_variants);
    if (!p->attr_done[vi]) {
        CUDA_TRY(cudaFuncSetAttribute(v->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v->smem));
        int occ = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v->fn, v->threads, v->smem));
        if (occ < 1) return fail(PSG_ERR_CUDA, "variant %s does not fit on an SM", v->name);
        p->occ[vi] = occ;
        p->attr_done[vi] = 1;
    }
    return PSG_OK;
}

static int launch_fused(psg_plan* p, const Variant* v, StiArgs a, int ncs, int frames_per_col, int min_iters,
                        cudaStream_t st, const Variant** used = nullptr) {
    const int vi = (int)(v - g_variants);
    const int N = 1 << v->logn;
    {
        int rcv = variant_ready(p, v);
        if (rcv) return rcv;
    }
    // geometry: gpc lanes per column inside a CTA, nsplit CTAs per column
    const int F = v->F;
    const int gpc = std::min(F, floor_pow2(frames_per_col));
    const int cpc = F / gpc;
    const int colblocks = (ncs + cpc - 1) / cpc;
    const int iters = (frames_per_col + gpc - 1) / gpc;  // per column if one CTA did it all
    const long long slots = (long long)p->sms * p->occ[vi];
    const long long target = slots * g_items_per_slot.load();
    int nsplit = (int)std::min<long long>((target + colblocks - 1) / colblocks, 1 << 20);
    // enough column blocks to keep every slot busy for three rounds: do not split columns at all (no
    // partial sums, no finalize launch; measured equal kernel time on cfg2)
    if (colblocks >= 3 * slots) nsplit = 1;
    // at most 1024 frames per fp32 accumulator (longer columns are split and summed in fp64)
    const int smin = (iters + 1023) / 1024, smax = std::max(1, iters / min_iters);
    nsplit = std::max(smin, std::min(nsplit, smax));
    nsplit = std::max(nsplit, 1);
    int chunk = ((iters + nsplit - 1) / nsplit) * gpc;
    nsplit = (frames_per_col + chunk - 1) / chunk;
    a.gpc = gpc;
    a.chunk = chunk;
    a.nsplit = nsplit;
    a.nfr = frames_per_col;
    if (nsplit > 1) {
        const size_t need = (size_t)ncs * nsplit * N;
        size_t have_b = p->partial_elems * sizeof(float);
        int rc = ensure_buffer((void**)&p->d_partial, &have_b, need * sizeof(float));
        p->partial_elems = have_b / sizeof(float);
        if (rc) return rc;
        a.partial = p->d_partial;
    }
    long long grid = (long long)colblocks * nsplit;
    if (grid > 2147483647ll) return fail(PSG_ERR_ARG, "psg_sti_run: grid too large");
    a.cb = 1;
    if (frames_per_col == 1 && !v->multi && g_use_multi.load()) {
        // Mode R: one frame per column -> the twin kernel that runs several column blocks per CTA
        const Variant* vm = variant_by_name((std::string(v->name) + "_m").c_str());
        if (vm && variant_ready(p, vm) == PSG_OK) {
            const long long slots_m = (long long)p->sms * p->occ[(int)(vm - g_variants)];
            int cb = 16;
            while (cb > 1 && (grid + cb - 1) / cb < slots_m * 4) cb /= 2;  // keep every slot busy for several rounds
            if (cb > 1) {
                a.cb = cb;
                grid = (grid + cb - 1) / cb;
                v = vm;
            }
        }
    } else if (v->multi) {
        if (frames_per_col != 1) return fail(PSG_ERR_ARG, "variant %s handles one frame per column only", v->name);
        a.cb = 4;
        grid = (grid + 3) / 4;
    }
    if (used) *used = v;
    void* args[] = {(void*)&a};
    CUDA_TRY(cudaLaunchKernel(v->fn, dim3((unsigned)grid), dim3(v->threads), args, v->smem, st));
    g_launches++;
    if (nsplit > 1) {
        const size_t total = (size_t)ncs * N;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)p->sms * 16);
        sti_finalize_kernel<<<blocks, 256, 0, st>>>(a.partial, nsplit, N, (size_t)ncs, a.scale, a.eps, a.out_lin,
                                                    a.out_db);
        g_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return PSG_OK;
}

// upload (once per plan) the per-pass twiddle tables of a variant
static int upload_pass_tables(const Variant* v, float2** d_out) {
    int r[4], np = variant_radices(v, r);
    std::vector<float2> t = build_pass_tables(1 << v->logn, r, np, v->twp);
    CUDA_TRY(cudaMalloc(d_out, sizeof(float2) * t.size()));
    CUDA_TRY(cudaMemcpy(*d_out, t.data(), sizeof(float2) * t.size(), cudaMemcpyHostToDevice));
    return PSG_OK;
}

// tables shared by the two large-nfft paths (nfft = r0 * 4096): first-pass twiddles W_N^{n'*k0},
// a window of ones and the pass tables of the 4096-point sub-transform
static int ensure_split_tables(psg_plan* p, const Variant* v) {
    constexpr int N2 = 4096;
    const int N = p->nfft, r0 = N / N2;
    if (p->d_twa) return PSG_OK;
    std::vector<float2> t((size_t)(r0 - 1) * N2);
    for (int k = 1; k < r0; ++k)
        for (int n = 0; n < N2; ++n) {
            const double ang = -2.0 * M_PI * (double)((long long)n * k) / (double)N;
            t[(size_t)(k - 1) * N2 + n] = make_float2((float)cos(ang), (float)sin(ang));
        }
    std::vector<float> ones(N2, 1.0f);
    CUDA_TRY(cudaMalloc(&p->d_twa, sizeof(float2) * t.size()));
    CUDA_TRY(cudaMemcpy(p->d_twa, t.data(), sizeof(float2) * t.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&p->d_ones, sizeof(float) * N2));
    CUDA_TRY(cudaMemcpy(p->d_ones, ones.data(), sizeof(float) * N2, cudaMemcpyHostToDevice));
    return upload_pass_tables(v, &p->d_twp_sub);
}

// ---- cluster path (sti_cluster.cuh): nfft = r0 * 4096 in one kernel, r0 CTAs per frame ----------
template <int R0, int ROWTMA>
static const void* cluster_fn_iq(int iqt) {
    return iqt == IQ_CI16 ? (const void*)sti_cluster_kernel<R0, IQ_CI16, ROWTMA>
           : iqt == IQ_CI8 ? (const void*)sti_cluster_kernel<R0, IQ_CI8, ROWTMA>
                           : (const void*)sti_cluster_kernel<R0, IQ_C64, ROWTMA>;
}
template <int ROWTMA>
static const void* cluster_fn_r0(int r0, int iqt) {
    switch (r0) {
        case 2: return cluster_fn_iq<2, ROWTMA>(iqt);
        case 4: return cluster_fn_iq<4, ROWTMA>(iqt);
        case 8: return cluster_fn_iq<8, ROWTMA>(iqt);
        case 16: return cluster_fn_iq<16, ROWTMA>(iqt);
    }
    return nullptr;
}
template <int R0>
static const void* dsmem_fn_iq(int iqt) {
    return iqt == IQ_CI16 ? (const void*)sti_dsmem_kernel<R0, IQ_CI16>
           : iqt == IQ_CI8 ? (const void*)sti_dsmem_kernel<R0, IQ_CI8>
                           : (const void*)sti_dsmem_kernel<R0, IQ_C64>;
}
static const void* dsmem_fn_r0(int r0, int iqt) {
    switch (r0) {
        case 2: return dsmem_fn_iq<2>(iqt);
        case 4: return dsmem_fn_iq<4>(iqt);
        case 8: return dsmem_fn_iq<8>(iqt);
        case 16: return dsmem_fn_iq<16>(iqt);
    }
    return nullptr;
}
static size_t dsmem_smem(int r0, int iqt) {
    const size_t iqb = iqt == IQ_C64 ? 8 : iqt == IQ_CI16 ? 4 : 2;
    const size_t seg = (size_t)(4096 / r0) * iqb + 16;
    return 192 + (size_t)r0 * seg + 2 * (size_t)(psg_pad(4096) + 2) * 8;
}
static size_t cluster_smem(int r0, int iqt) {
    const size_t iqb = iqt == IQ_C64 ? 8 : iqt == IQ_CI16 ? 4 : 2;
    const size_t seg = (size_t)(4096 / r0) * iqb + 16;
    return 128 + 4 * 256 * 8 + 2 * (size_t)r0 * seg + (size_t)(psg_pad(4096) + 2) * 8;
}

// returns PSG_OK with *ran = false when the device cannot co-schedule the cluster (caller falls back)
static int run_cluster(psg_plan* p, const StiArgs& a, int ncs, int frames_per_col, int rowtma, cudaStream_t st, bool* ran) {
    constexpr int N2 = 4096;
    const int N = p->nfft, r0 = N / N2;
    *ran = false;
    const Variant* v = variant_by_name(g_default_tma[12 - 5]);
    if (!v || v->twp != 2) return fail(PSG_ERR_UNSUPPORTED, "cluster path needs the power-layout 4096-point tables");
    // rowtma: 0 = exchange through L2, rows loaded to registers; 1 = L2, rows by bulk copy; 2 = DSMEM
    const void* fn = rowtma == 2 ? dsmem_fn_r0(r0, a.iq_type) : rowtma ? cluster_fn_r0<1>(r0, a.iq_type) : cluster_fn_r0<0>(r0, a.iq_type);
    if (!fn) return fail(PSG_ERR_UNSUPPORTED, "cluster path: r0=%d", r0);
    const size_t smem = rowtma == 2 ? dsmem_smem(r0, a.iq_type) : cluster_smem(r0, a.iq_type);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)r0;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // occupancy depends on the element type only through the stage size; query per call signature once
    static thread_local const void* q_fn = nullptr;
    static thread_local int q_dev = -1, q_slots = 0;
    if (q_fn != fn || q_dev != p->device) {
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (r0 > 8) CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cfg.gridDim = dim3((unsigned)(r0 * p->sms * 2));
        int nmax = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nmax, fn, &cfg);
        if (e != cudaSuccess) { cudaGetLastError(); nmax = 0; }
        q_fn = fn;
        q_dev = p->device;
        q_slots = nmax;
    }
    if (q_slots < 1) return PSG_OK;  // not schedulable here: caller uses the split path
    int rc = ensure_split_tables(p, v);
    if (rc) return rc;
    // items: whole columns when there are enough of them, else columns split into frame chunks
    // (>= 4 frames each: the pipeline is two frames deep; <= 1024 frames per fp32 accumulator)
    const long long want = 16ll * q_slots;
    int nsplit = 1;
    if (ncs < want) nsplit = (int)std::min<long long>((want + ncs - 1) / ncs, std::max(1, frames_per_col / 4));
    nsplit = std::max(nsplit, (frames_per_col + 1023) / 1024);
    const int chunk = (frames_per_col + nsplit - 1) / nsplit;
    nsplit = (frames_per_col + chunk - 1) / chunk;
    const long long nitems = (long long)ncs * nsplit;
    const int nclusters = (int)std::min<long long>(q_slots, nitems);
    if (rowtma != 2) {
        rc = ensure_buffer((void**)&p->d_xslot, &p->xslot_bytes, (size_t)q_slots * 3 * N * 8);
        if (rc) return rc;
    }
    rc = ensure_buffer((void**)&p->d_tmp, &p->tmp_bytes, (size_t)nitems * N * 4);
    if (rc) return rc;
    ClusterArgs ca;
    ca.iq = a.iq;
    ca.sub_stride = a.sub_stride;
    ca.hop_elems = a.hop_elems;
    ca.col_off = a.col_off;
    ca.ncol = a.ncol;
    ca.ncs = ncs;
    ca.nfr = frames_per_col;
    ca.chunk = chunk;
    ca.nsplit = nsplit;
    ca.nclusters = nclusters;
    ca.win = p->d_win;
    ca.twa = p->d_twa;
    ca.twp = p->d_twp_sub;
    ca.scratch = p->d_xslot;
    ca.tmp = p->d_tmp;
    cfg.gridDim = dim3((unsigned)(nclusters * r0));
    if (getenv("PSG_DEBUG"))
        fprintf(stderr, "[psg] cluster path r0=%d resident clusters=%d items=%lld (nsplit=%d, chunk=%d) smem=%zu\n", r0, q_slots,
                nitems, nsplit, chunk, smem);
    void* args[] = {(void*)&ca};
    CUDA_TRY(cudaLaunchKernelExC(&cfg, fn, args));
    g_launches++;
    ClusterFinArgs fa;
    fa.tmp = p->d_tmp;
    fa.r0 = r0;
    fa.nsplit = nsplit;
    fa.scale = a.scale;
    fa.eps = a.eps;
    fa.out_lin = a.out_lin;
    fa.out_db = a.out_db;
    sti_cluster_finalize_kernel<<<dim3(N2 / 128, (unsigned)ncs), 256, 0, st>>>(fa);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    snprintf(p->variant_name, sizeof(p->variant_name), "cluster%dx4096%s%s", r0, rowtma == 2 ? "_dsmem" : rowtma ? "_tma" : "_ldg",
             a.iq_type == IQ_CI16 ? "_i16" : a.iq_type == IQ_CI8 ? "_i8" : "");
    *ran = true;
    return PSG_OK;
}

// ---- whole-frame path (sti_whole.cuh): nfft = 8192 / 16384 with the frame resident in one SM -----
template <int R0, int NST>
static const void* whole_fn_iq(int iqt) {
    return iqt == IQ_CI16 ? (const void*)sti_whole_kernel<R0, IQ_CI16, NST>
           : iqt == IQ_CI8 ? (const void*)sti_whole_kernel<R0, IQ_CI8, NST>
                           : (const void*)sti_whole_kernel<R0, IQ_C64, NST>;
}
static size_t whole_smem(int r0, int iqt, int nst) {
    const size_t iqb = iqt == IQ_C64 ? 8 : iqt == IQ_CI16 ? 4 : 2;
    return 128 + (size_t)nst * r0 * (512 * iqb + 16) + (size_t)(psg_pad(r0 * 4096) + 2) * 8;
}

static void fill_whole_constants(WholeArgs& wa, int N) {
    for (int q = 0; q < 4; ++q)
        for (int j = 0; j < 16; ++j) {
            const double ang = -2.0 * M_PI * (double)((256ll * j << q) % N) / (double)N;
            wa.cm[q][j] = make_float2((float)cos(ang), (float)sin(ang));
        }
}

// items of the whole-frame kernels: whole columns, or columns split into frame chunks (>= 8 frames each,
// <= 1024 per fp32 accumulator) when that fills the last wave of `slots` concurrent items better
static int whole_nsplit(int ncs, int frames_per_col, long long slots) {
    const int smin = (frames_per_col + 1023) / 1024;
    const int smax = std::max(smin, frames_per_col / 8);
    int best = smin;
    double best_eff = 0.0;
    for (int ns = smin; ns <= std::min(smax, smin + 15); ++ns) {
        const long long items = (long long)ncs * ns;
        const long long waves = (items + slots - 1) / slots;
        const double eff = (double)items / (double)(waves * slots) - 0.01 * (ns - smin);  // mild preference for fewer splits
        if (eff > best_eff + 1e-9) { best_eff = eff; best = ns; }
    }
    return best;
}

static int run_whole(psg_plan* p, const StiArgs& a0, int ncs, int frames_per_col, int nst, cudaStream_t st) {
    constexpr int N2 = 4096;
    const int N = p->nfft, r0 = N / N2;
    const Variant* v = variant_by_name(g_default_tma[12 - 5]);
    if (!v || v->twp != 2) return fail(PSG_ERR_UNSUPPORTED, "whole-frame path needs the power-layout 4096-point tables");
    const void* fn = nullptr;
    if (r0 == 4 && nst == 4) fn = whole_fn_iq<4, 4>(a0.iq_type);
    else if (r0 == 4 && nst == 2) fn = whole_fn_iq<4, 2>(a0.iq_type);
    else if (r0 == 2 && nst == 4) fn = whole_fn_iq<2, 4>(a0.iq_type);
    else if (r0 == 2 && nst == 8) fn = whole_fn_iq<2, 8>(a0.iq_type);
    if (!fn) return fail(PSG_ERR_UNSUPPORTED, "whole-frame path: nfft=%d stages=%d", N, nst);
    const size_t smem = whole_smem(r0, a0.iq_type, nst);
    static thread_local const void* q_fn = nullptr;
    static thread_local int q_dev = -1;
    if (q_fn != fn || q_dev != p->device) {
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        q_fn = fn;
        q_dev = p->device;
    }
    int rc = ensure_split_tables(p, v);
    if (rc) return rc;
    int nsplit = whole_nsplit(ncs, frames_per_col, p->sms);
    const int chunk = (frames_per_col + nsplit - 1) / nsplit;
    nsplit = (frames_per_col + chunk - 1) / chunk;
    WholeArgs wa;
    wa.s = a0;
    wa.s.twp = p->d_twp_sub;
    wa.s.gpc = 1;
    wa.s.chunk = chunk;
    wa.s.nsplit = nsplit;
    wa.s.nfr = frames_per_col;
    fill_whole_constants(wa, N);
    if (nsplit > 1) {
        size_t have_b = p->partial_elems * sizeof(float);
        rc = ensure_buffer((void**)&p->d_partial, &have_b, (size_t)ncs * nsplit * N * sizeof(float));
        p->partial_elems = have_b / sizeof(float);
        if (rc) return rc;
        wa.s.partial = p->d_partial;
    }
    const long long grid = (long long)ncs * nsplit;
    if (grid > 2147483647ll) return fail(PSG_ERR_ARG, "psg_sti_run: grid too large");
    void* args[] = {(void*)&wa};
    CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(512), args, smem, st));
    g_launches++;
    if (nsplit > 1) {
        const size_t total = (size_t)ncs * N;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)p->sms * 16);
        sti_finalize_kernel<<<blocks, 256, 0, st>>>(wa.s.partial, nsplit, N, (size_t)ncs, wa.s.scale, wa.s.eps, wa.s.out_lin,
                                                    wa.s.out_db);
        g_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    snprintf(p->variant_name, sizeof(p->variant_name), "whole%dx4096_s%d%s", r0, nst,
             a0.iq_type == IQ_CI16 ? "_i16" : a0.iq_type == IQ_CI8 ? "_i8" : "");
    return PSG_OK;
}

// ---- whole-frame path with both radix-2 passes in registers (sti_whole16.cuh): nfft = 16384 ---------------
static int run_whole16(psg_plan* p, const StiArgs& a0, int ncs, int frames_per_col, cudaStream_t st) {
    const int N = p->nfft;
    if (N != 16384) return fail(PSG_ERR_UNSUPPORTED, "whole_f path: nfft=%d (16384 only)", N);
    const void* fn = a0.iq_type == IQ_CI16 ? (const void*)sti_whole16_kernel<IQ_CI16>
                     : a0.iq_type == IQ_CI8 ? (const void*)sti_whole16_kernel<IQ_CI8>
                                            : (const void*)sti_whole16_kernel<IQ_C64>;
    const size_t smem = a0.iq_type == IQ_CI16 ? Whole16Cfg<IQ_CI16>::smem_bytes
                        : a0.iq_type == IQ_CI8 ? Whole16Cfg<IQ_CI8>::smem_bytes
                                               : Whole16Cfg<IQ_C64>::smem_bytes;
    static thread_local const void* q_fn = nullptr;
    static thread_local int q_dev = -1;
    if (q_fn != fn || q_dev != p->device) {
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        q_fn = fn;
        q_dev = p->device;
    }
    int nsplit = whole_nsplit(ncs, frames_per_col, p->sms);
    const int chunk = (frames_per_col + nsplit - 1) / nsplit;
    nsplit = (frames_per_col + chunk - 1) / chunk;
    Whole16Args wa;
    wa.s = a0;
    wa.s.gpc = 1;
    wa.s.chunk = chunk;
    wa.s.nsplit = nsplit;
    wa.s.nfr = frames_per_col;
    if (nsplit > 1) {
        size_t have_b = p->partial_elems * sizeof(float);
        int rc = ensure_buffer((void**)&p->d_partial, &have_b, (size_t)ncs * nsplit * N * sizeof(float));
        p->partial_elems = have_b / sizeof(float);
        if (rc) return rc;
        wa.s.partial = p->d_partial;
    }
    const long long grid = (long long)ncs * nsplit;
    if (grid > 2147483647ll) return fail(PSG_ERR_ARG, "psg_sti_run: grid too large");
    void* args[] = {(void*)&wa};
    CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(512), args, smem, st));
    g_launches++;
    if (nsplit > 1) {
        const size_t total = (size_t)ncs * N;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)p->sms * 16);
        sti_finalize_kernel<<<blocks, 256, 0, st>>>(wa.s.partial, nsplit, N, (size_t)ncs, wa.s.scale, wa.s.eps, wa.s.out_lin,
                                                    wa.s.out_db);
        g_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    snprintf(p->variant_name, sizeof(p->variant_name), "whole16x2x16x2x16%s",
             a0.iq_type == IQ_CI16 ? "_i16" : a0.iq_type == IQ_CI8 ? "_i8" : "");
    return PSG_OK;
}

// ---- clustered whole-frame path (sti_whole.cuh): nfft = 32768 / 65536 on clusters of 2 / 4 CTAs ---------
template <int CL, int ROWS>
static const void* wholec_fn_iq(int iqt) {
    return iqt == IQ_CI16 ? (const void*)sti_wholec_kernel<CL, ROWS, IQ_CI16>
           : iqt == IQ_CI8 ? (const void*)sti_wholec_kernel<CL, ROWS, IQ_CI8>
                           : (const void*)sti_wholec_kernel<CL, ROWS, IQ_C64>;
}

// rows = 4: 512 threads, one CTA per SM (32768 on 2 CTAs, 65536 on 4); rows = 2: 256 threads, two CTAs per SM
// (16384 on 2 CTAs, 32768 on 4, 65536 on 8).  Returns PSG_OK with *ran = false when the device cannot
// co-schedule the cluster (caller falls back)
static int run_wholec(psg_plan* p, const StiArgs& a0, int ncs, int frames_per_col, int rows, cudaStream_t st, bool* ran) {
    constexpr int N2 = 4096;
    const int N = p->nfft, cl = N / (rows * N2);
    *ran = false;
    const Variant* v = variant_by_name(g_default_tma[12 - 5]);
    if (!v || v->twp != 2) return fail(PSG_ERR_UNSUPPORTED, "whole-frame path needs the power-layout 4096-point tables");
    const void* fn = nullptr;
    if (rows == 4 && cl == 2) fn = wholec_fn_iq<2, 4>(a0.iq_type);
    else if (rows == 4 && cl == 4) fn = wholec_fn_iq<4, 4>(a0.iq_type);
    else if (rows == 2 && cl == 2) fn = wholec_fn_iq<2, 2>(a0.iq_type);
    else if (rows == 2 && cl == 4) fn = wholec_fn_iq<4, 2>(a0.iq_type);
    else if (rows == 2 && cl == 8) fn = wholec_fn_iq<8, 2>(a0.iq_type);
    if (!fn) return fail(PSG_ERR_UNSUPPORTED, "clustered whole-frame path: nfft=%d rows=%d", N, rows);
    const size_t iqb = a0.iq_type == IQ_C64 ? 8 : a0.iq_type == IQ_CI16 ? 4 : 2;
    const int threads = 128 * rows;
    const size_t smem = 128 + 4 * 4 * ((size_t)threads * iqb + 16) + (size_t)(psg_pad(rows * N2) + 2) * 8;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static thread_local const void* q_fn = nullptr;
    static thread_local int q_dev = -1, q_slots = 0;
    if (q_fn != fn || q_dev != p->device) {
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cfg.gridDim = dim3((unsigned)(cl * p->sms * 2));
        int nmax = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nmax, fn, &cfg);
        if (e != cudaSuccess) { cudaGetLastError(); nmax = 0; }
        q_fn = fn;
        q_dev = p->device;
        q_slots = nmax;
    }
    if (q_slots < 1) return PSG_OK;
    int rc = ensure_split_tables(p, v);
    if (rc) return rc;
    int nsplit = whole_nsplit(ncs, frames_per_col, q_slots);
    const int chunk = (frames_per_col + nsplit - 1) / nsplit;
    nsplit = (frames_per_col + chunk - 1) / chunk;
    WholeArgs wa;
    wa.s = a0;
    wa.s.twp = p->d_twp_sub;
    wa.s.gpc = 1;
    wa.s.chunk = chunk;
    wa.s.nsplit = nsplit;
    wa.s.nfr = frames_per_col;
    fill_whole_constants(wa, N);
    if (nsplit > 1) {
        size_t have_b = p->partial_elems * sizeof(float);
        rc = ensure_buffer((void**)&p->d_partial, &have_b, (size_t)ncs * nsplit * N * sizeof(float));
        p->partial_elems = have_b / sizeof(float);
        if (rc) return rc;
        wa.s.partial = p->d_partial;
    }
    const long long grid = (long long)ncs * nsplit * cl;
    if (grid > 2147483647ll) return fail(PSG_ERR_ARG, "psg_sti_run: grid too large");
    cfg.gridDim = dim3((unsigned)grid);
    if (getenv("PSG_DEBUG"))
        fprintf(stderr, "[psg] clustered whole-frame path cl=%d rows=%d resident clusters=%d items=%lld (nsplit=%d, chunk=%d)\n", cl,
                rows, q_slots, (long long)ncs * nsplit, nsplit, chunk);
    void* args[] = {(void*)&wa};
    CUDA_TRY(cudaLaunchKernelExC(&cfg, fn, args));
    g_launches++;
    if (nsplit > 1) {
        const size_t total = (size_t)ncs * N;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)p->sms * 16);
        sti_finalize_kernel<<<blocks, 256, 0, st>>>(wa.s.partial, nsplit, N, (size_t)ncs, wa.s.scale, wa.s.eps, wa.s.out_lin,
                                                    wa.s.out_db);
        g_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    snprintf(p->variant_name, sizeof(p->variant_name), "whole%dx4096_c%d%s%s", rows * cl, cl, rows == 2 ? "r2" : "",
             a0.iq_type == IQ_CI16 ? "_i16" : a0.iq_type == IQ_CI8 ? "_i8" : "");
    *ran = true;
    return PSG_OK;
}

// ---- radix-32 whole-frame path (sti_r32.cuh, psg_r32.cu): nfft = 16384 / 32768 / 65536, three shared-memory passes ----
// Persistent CTAs (clusters of 2 / 4 for 32768 / 65536) walk the work items; *ran = false when the device cannot
// co-schedule the cluster (caller falls back).
static int run_r32(psg_plan* p, const StiArgs& a0, int ncs, int frames_per_col, cudaStream_t st, bool* ran) {
    const int N = p->nfft;
    *ran = false;
    static thread_local int q_key = -1, q_groups = 0;
    const int key = (p->device << 8) | (p->logn << 2) | a0.iq_type;
    if (q_key != key) {
        int ng = 0;
        const cudaError_t e = (cudaError_t)psg_r32_max_groups(p->logn, a0.iq_type, p->device, p->sms, &ng);
        if (e != cudaSuccess) return fail(PSG_ERR_CUDA, "radix-32 path: %s", cudaGetErrorString(e));
        q_key = key;
        q_groups = ng;
    }
    if (q_groups < 1) return PSG_OK;
    int nsplit = whole_nsplit(ncs, frames_per_col, q_groups);
    const int chunk = (frames_per_col + nsplit - 1) / nsplit;
    nsplit = (frames_per_col + chunk - 1) / chunk;
    StiArgs a = a0;
    a.gpc = 1;
    a.chunk = chunk;
    a.nsplit = nsplit;
    a.nfr = frames_per_col;
    if (nsplit > 1) {
        size_t have_b = p->partial_elems * sizeof(float);
        const int rc = ensure_buffer((void**)&p->d_partial, &have_b, (size_t)ncs * nsplit * N * sizeof(float));
        p->partial_elems = have_b / sizeof(float);
        if (rc) return rc;
        a.partial = p->d_partial;
    }
    const long long nitems = (long long)ncs * nsplit;
    if (nitems > 2147483647ll) return fail(PSG_ERR_ARG, "psg_sti_run: too many work items");
    const int ngroups = (int)std::min<long long>(nitems, q_groups);
    if (getenv("PSG_DEBUG"))
        fprintf(stderr, "[psg] radix-32 path nfft=%d groups=%d (resident %d) items=%lld (nsplit=%d, chunk=%d)\n", N, ngroups, q_groups,
                nitems, nsplit, chunk);
    const cudaError_t e = (cudaError_t)psg_r32_launch(p->logn, a0.iq_type, a, (int)nitems, ngroups, st);
    if (e != cudaSuccess) return fail(PSG_ERR_CUDA, "radix-32 kernel launch: %s", cudaGetErrorString(e));
    g_launches++;
    if (nsplit > 1) {
        const size_t total = (size_t)ncs * N;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)p->sms * 16);
        sti_finalize_kernel<<<blocks, 256, 0, st>>>(a.partial, nsplit, N, (size_t)ncs, a.scale, a.eps, a.out_lin, a.out_db);
        g_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    int thr = 0, cl = 0;
    psg_r32_describe(p->logn, &thr, &cl);
    snprintf(p->variant_name, sizeof(p->variant_name), "r32_%s_t%dc%d%s",
             N == 8192 ? "32x16x16" : N == 16384 ? "32x32x16" : N == 32768 ? "32x32x32" : "32x32x2x32", thr, cl,
             a0.iq_type == IQ_CI16 ? "_i16" : a0.iq_type == IQ_CI8 ? "_i8" : "");
    *ran = true;
    return PSG_OK;
}

// nfft = r0 * 4096 (r0 = 2..16): streaming first pass -> L2-resident scratch -> tuned 4096-point
// fused kernel over the r0 sub-sequences -> interleave.  See sti_kernels.cuh.
static int run_split(psg_plan* p, StiArgs a, int ncs, int frames_per_col, cudaStream_t st) {
    constexpr int N2 = 4096;
    const int N = p->nfft, r0 = N / N2;
    const Variant* v = variant_by_name(g_default_tma[12 - 5]);
    if (!v) return fail(PSG_ERR_UNSUPPORTED, "no 4096-point variant for the split path");
    {
        int rct = ensure_split_tables(p, v);
        if (rct) return rct;
    }
    snprintf(p->variant_name, sizeof(p->variant_name), "split%dx4096+%s", r0, v->name);
    // chunking: whole columns per chunk when a column's frames fit the scratch, else one column at a
    // time in frame blocks whose raw sums are carried in d_carry
    const long long cap_frames = std::max<long long>(1, (long long)((size_t)g_split_scratch_bytes.load() / ((size_t)N * 8)));
    const int cols_per_chunk = (int)std::max<long long>(1, std::min<long long>(ncs, cap_frames / frames_per_col));
    const int frames_per_block = (int)std::min<long long>(frames_per_col, cap_frames);
    const bool carried = frames_per_block < frames_per_col;
    int rc = ensure_buffer((void**)&p->d_scratch, &p->scratch_bytes, (size_t)cols_per_chunk * frames_per_block * N * 8);
    if (rc) return rc;
    rc = ensure_buffer((void**)&p->d_tmp, &p->tmp_bytes, (size_t)cols_per_chunk * N * 4);
    if (rc) return rc;
    if (carried) {
        rc = ensure_buffer((void**)&p->d_carry, &p->carry_bytes, (size_t)cols_per_chunk * N * 4);
        if (rc) return rc;
    }
    // column c of a chunk starts at frame c * frames_per_block of the scratch (fixed stride, so the
    // offset table phase B reads is uploaded once per geometry)
    if (p->h_colb.size() != (size_t)cols_per_chunk || p->colb_stride != (long long)frames_per_block * N) {
        rc = ensure_buffer((void**)&p->d_colb, &p->colb_bytes, (size_t)cols_per_chunk * 8);
        if (rc) return rc;
        p->colb_stride = (long long)frames_per_block * N;
        p->h_colb.resize(cols_per_chunk);
        for (int c = 0; c < cols_per_chunk; ++c) p->h_colb[c] = (long long)c * p->colb_stride;
        CUDA_TRY(cudaMemcpyAsync(p->d_colb, p->h_colb.data(), (size_t)cols_per_chunk * 8, cudaMemcpyHostToDevice, st));
    }
    const float scale = a.scale, eps = a.eps;
    float* out_lin = a.out_lin;
    float* out_db = a.out_db;
    for (int cs_lo = 0; cs_lo < ncs; cs_lo += cols_per_chunk) {
        const int nc = std::min(cols_per_chunk, ncs - cs_lo);
        for (int k_lo = 0; k_lo < frames_per_col; k_lo += frames_per_block) {
            const int nf = std::min(frames_per_block, frames_per_col - k_lo);
            // phase A
            SplitArgs sa;
            sa.iq = a.iq;
            sa.iq_type = a.iq_type;
            sa.sample_stride = a.sample_stride;
            sa.sub_stride = a.sub_stride;
            sa.hop_elems = a.hop_elems;
            sa.col_off = a.col_off;
            sa.ncol = a.ncol;
            sa.cs_lo = cs_lo;
            sa.ncs_chunk = nc;
            sa.k_lo = k_lo;
            sa.nfr_chunk = nf;
            sa.col_frames = frames_per_block;
            sa.win = p->d_win;
            sa.twa = p->d_twa;
            sa.scratch = p->d_scratch;
            constexpr int FPB = 4;
            const dim3 ga(N2 / 256, (unsigned)((nc * (long long)nf + FPB - 1) / FPB));
            switch (r0) {
                case 2: sti_split_pass_kernel<2, FPB><<<ga, 256, 0, st>>>(sa); break;
                case 4: sti_split_pass_kernel<4, FPB><<<ga, 256, 0, st>>>(sa); break;
                case 8: sti_split_pass_kernel<8, FPB><<<ga, 256, 0, st>>>(sa); break;
                case 16: sti_split_pass_kernel<16, FPB><<<ga, 256, 0, st>>>(sa); break;
                default: return fail(PSG_ERR_UNSUPPORTED, "split path: r0=%d", r0);
            }
            g_launches++;
            // phase B: r0 * nc virtual columns of nf frames each
            StiArgs b;
            b.iq = p->d_scratch;
            b.iq_type = IQ_C64;
            b.sample_stride = 1;
            b.sub_stride = N2;
            b.hop_elems = N;
            b.col_off = p->d_colb;
            b.ncol = nc;
            b.nsub = r0;
            b.win = p->d_ones;
            b.tw = nullptr;
            b.twp = p->d_twp_sub;
            b.scale = 1.0f;
            b.eps = eps;
            b.out_lin = p->d_tmp;
            b.out_db = nullptr;
            b.partial = nullptr;
            rc = launch_fused(p, v, b, nc * r0, nf, 8, st);
            if (rc) return rc;
            // phase C
            InterleaveArgs ia;
            ia.tmp = p->d_tmp;
            ia.r0 = r0;
            ia.ncs_chunk = nc;
            ia.cs_lo = cs_lo;
            ia.first = k_lo == 0;
            ia.last = k_lo + nf >= frames_per_col;
            ia.scale = scale;
            ia.eps = eps;
            ia.acc = carried ? p->d_carry : nullptr;
            ia.out_lin = out_lin;
            ia.out_db = out_db;
            sti_interleave_kernel<<<dim3(N2 / 128, nc), 256, 0, st>>>(ia);
            g_launches++;
        }
    }
    CUDA_TRY(cudaGetLastError());
    return PSG_OK;
}

// nfft = 2^a 3^b 5^c: direct mixed-radix transform in shared memory (sti_mixed_kernel)
// Round lengths with a compile-time plan (sti_mixct.cuh, mixct_plans.inc): same work split, kernel from psg_mixct.cu
static int run_mixct(psg_plan* p, StiArgs a, int ncs, int frames_per_col, const MixctInfo& mi, cudaStream_t st) {
    const int N = p->nfft;
    if (mi.occ < 1) return fail(PSG_ERR_CUDA, "mixed-radix kernel (nfft=%d) does not fit on an SM", N);
    const long long slots = (long long)p->sms * mi.occ;
    // items: at most ~1024 frames each; enough of them to fill the resident CTAs a few times over while a CTA's
    // groups still have several frames each
    int nsplit = std::max(1, (frames_per_col + 1023) / 1024);
    if ((long long)ncs * nsplit < 4 * slots)
        nsplit = (int)std::max<long long>(nsplit, std::min<long long>(std::max(1, frames_per_col / (8 * mi.groups)), (4 * slots + ncs - 1) / ncs));
    const int chunk = (frames_per_col + nsplit - 1) / nsplit;
    nsplit = (frames_per_col + chunk - 1) / chunk;
    a.gpc = 1;
    a.chunk = chunk;
    a.nsplit = nsplit;
    if (nsplit > 1) {
        size_t have_b = p->partial_elems * sizeof(float);
        int rc = ensure_buffer((void**)&p->d_partial, &have_b, (size_t)ncs * nsplit * N * sizeof(float));
        p->partial_elems = have_b / sizeof(float);
        if (rc) return rc;
        a.partial = p->d_partial;
    }
    const long long items = (long long)ncs * nsplit;
    const long long grid = std::min<long long>(items, slots);
    a.tw = p->d_tw;
    const int e = psg_mixct_launch(N, a.iq_type, frames_per_col, a, grid, st);
    if (e) return fail(PSG_ERR_CUDA, "mixed-radix kernel launch (nfft=%d): %s", N, cudaGetErrorString((cudaError_t)e));
    g_launches++;
    if (nsplit > 1) {
        const size_t total = (size_t)ncs * N;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)p->sms * 16);
        sti_finalize_kernel<<<blocks, 256, 0, st>>>(a.partial, nsplit, N, (size_t)ncs, a.scale, a.eps, a.out_lin, a.out_db);
        g_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    snprintf(p->variant_name, sizeof(p->variant_name), "%s", mi.name);
    return PSG_OK;
}

static int run_mixed(psg_plan* p, StiArgs a, int ncs, int frames_per_col, cudaStream_t st) {
    const int N = p->nfft;
    {
        bool force_rt = false;
        {
            std::lock_guard<NoMutex> lk(g_variant_mu);
            force_rt = g_variant_override == "mixed_rt";
        }
        MixctInfo mi;
        if (!force_rt && psg_mixct_query(N, a.iq_type, frames_per_col, &mi) == (int)cudaSuccess)
            return run_mixct(p, a, ncs, frames_per_col, mi, st);
        cudaGetLastError();
    }
    const int tpf = std::min(512, std::max(32, ((N / 8 + 31) / 32) * 32));
    const size_t group_bytes = (size_t)((psg_pad(N) + 3) & ~1) * 8 + (size_t)((N + 3) & ~3) * 4;
    int groups = std::max(1, (N >= 4096 ? 512 : 256) / tpf);
    while (groups > 1 && groups * group_bytes > 100 * 1024) groups /= 2;
    groups = std::min(groups, std::max(1, floor_pow2(frames_per_col)));
    const int threads = tpf * groups;
    const size_t smem = groups * group_bytes;
    const void* fn = (const void*)sti_mixed_kernel;
    static thread_local int q_dev = -1;
    if (q_dev != p->device) {
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        q_dev = p->device;
    }
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem));
    if (occ < 1) return fail(PSG_ERR_CUDA, "mixed-radix kernel (nfft=%d) does not fit on an SM", N);
    const long long slots = (long long)p->sms * occ;
    int nsplit = std::max(1, (frames_per_col + 1023) / 1024);
    if ((long long)ncs * nsplit < 2 * slots)
        nsplit = (int)std::max<long long>(nsplit, std::min<long long>(std::max(1, frames_per_col / (4 * groups)), (2 * slots + ncs - 1) / ncs));
    const int chunk = (frames_per_col + nsplit - 1) / nsplit;
    nsplit = (frames_per_col + chunk - 1) / chunk;
    a.gpc = 1;
    a.chunk = chunk;
    a.nsplit = nsplit;
    if (nsplit > 1) {
        size_t have_b = p->partial_elems * sizeof(float);
        int rc = ensure_buffer((void**)&p->d_partial, &have_b, (size_t)ncs * nsplit * N * sizeof(float));
        p->partial_elems = have_b / sizeof(float);
        if (rc) return rc;
        a.partial = p->d_partial;
    }
    const long long items = (long long)ncs * nsplit;
    const long long grid = std::min<long long>(items, slots);
    MixedArgs b;
    b.twf = p->d_tw;
    b.n = N;
    b.tpf = tpf;
    b.npass = p->mx_npass;
    for (int i = 0; i < 10; ++i) b.radix[i] = p->mx_radix[i];
    void* args[] = {(void*)&a, (void*)&b};
    CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(threads), args, smem, st));
    g_launches++;
    if (nsplit > 1) {
        const size_t total = (size_t)ncs * N;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)p->sms * 16);
        sti_finalize_kernel<<<blocks, 256, 0, st>>>(a.partial, nsplit, N, (size_t)ncs, a.scale, a.eps, a.out_lin, a.out_db);
        g_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    std::string name = "mixed" + std::to_string(N) + "_";
    for (int i = 0; i < p->mx_npass; ++i) name += (i ? "x" : "") + std::to_string(p->mx_radix[i]);
    snprintf(p->variant_name, sizeof(p->variant_name), "%s", name.c_str());
    return PSG_OK;
}

// non power-of-two nfft, M <= 16384: Bluestein with mixed-radix passes in shared memory (sti_bluestein.cuh)
static int run_bluestein16(psg_plan* p, StiArgs a, int ncs, int frames_per_col, cudaStream_t st) {
    const int N = p->nfft;
    const size_t M = (size_t)1 << p->logm;
    // frame groups of tpf threads (one radix-16 butterfly per thread and pass), as many as fit 256 threads
    // (512 for the two largest M) and half the shared memory of an SM
    const int tpf = (int)std::min<size_t>(512, std::max<size_t>(16, M / 16));
    const size_t group_bytes = (size_t)(psg_pad((int)M) + 2) * 8 + (size_t)((N + 3) & ~3) * 4;
    int groups = std::max(1, (M >= 8192 ? 512 : 256) / tpf);
    while (groups > 1 && groups * group_bytes > 100 * 1024) groups /= 2;
    groups = std::min(groups, std::max(1, floor_pow2(frames_per_col)));
    const int threads = tpf * groups;
    const size_t smem = groups * group_bytes;
    const void* fn = (const void*)sti_bluestein16_kernel;
    static thread_local int q_dev = -1;
    if (q_dev != p->device) {
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        q_dev = p->device;
    }
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem));
    if (occ < 1) return fail(PSG_ERR_CUDA, "bluestein16 (M=%zu) does not fit on an SM", M);
    const long long slots = (long long)p->sms * occ;
    // items: columns, split into frame chunks (>= 4 frames, <= 1024 per fp32 accumulator) when there are too few
    int nsplit = std::max(1, (frames_per_col + 1023) / 1024);
    if ((long long)ncs * nsplit < 2 * slots)
        nsplit = (int)std::max<long long>(nsplit, std::min<long long>(std::max(1, frames_per_col / (4 * groups)), (2 * slots + ncs - 1) / ncs));
    const int chunk = (frames_per_col + nsplit - 1) / nsplit;
    nsplit = (frames_per_col + chunk - 1) / chunk;
    a.gpc = 1;
    a.chunk = chunk;
    a.nsplit = nsplit;
    if (nsplit > 1) {
        size_t have_b = p->partial_elems * sizeof(float);
        int rc = ensure_buffer((void**)&p->d_partial, &have_b, (size_t)ncs * nsplit * N * sizeof(float));
        p->partial_elems = have_b / sizeof(float);
        if (rc) return rc;
        a.partial = p->d_partial;
    }
    const long long items = (long long)ncs * nsplit;
    const long long grid = std::min<long long>(items, slots);
    Bluestein16Args b;
    b.aw = p->d_aw;
    b.bpos = p->d_bpos;
    b.twf = p->d_twf;
    b.n = N;
    b.logm = p->logm;
    b.tpf = tpf;
    b.npass = p->bs_npass;
    for (int i = 0; i < 4; ++i) b.radix[i] = p->bs_radix[i];
    void* args[] = {(void*)&a, (void*)&b};
    CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(threads), args, smem, st));
    g_launches++;
    if (nsplit > 1) {
        const size_t total = (size_t)ncs * N;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)p->sms * 16);
        sti_finalize_kernel<<<blocks, 256, 0, st>>>(a.partial, nsplit, N, (size_t)ncs, a.scale, a.eps, a.out_lin, a.out_db);
        g_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    std::string name = "bluestein_m" + std::to_string(M) + "_";
    for (int i = 0; i < p->bs_npass; ++i) name += (i ? "x" : "") + std::to_string(p->bs_radix[i]);
    snprintf(p->variant_name, sizeof(p->variant_name), "%s", name.c_str());
    return PSG_OK;
}

// non power-of-two nfft: radix-2 Bluestein kernel (M > 16384: work buffer in global scratch; or forced)
static int run_bluestein(psg_plan* p, StiArgs a, int ncs, int frames_per_col, cudaStream_t st) {
    const int N = p->nfft;
    const size_t M = (size_t)1 << p->logm;
    {
        bool force_r2 = false, force_bs = false;
        {
            std::lock_guard<NoMutex> lk(g_variant_mu);
            force_r2 = g_variant_override == "bluestein_r2";
            force_bs = g_variant_override == "bluestein";
        }
        if (p->mx_npass && !force_r2 && !force_bs) return run_mixed(p, a, ncs, frames_per_col, st);
        if (p->d_bpos && !force_r2) return run_bluestein16(p, a, ncs, frames_per_col, st);
    }
    snprintf(p->variant_name, sizeof(p->variant_name), "bluestein_m%zu_r2", M);
    const size_t smem_need = M * 8 + (size_t)N * 4;
    const bool in_smem = smem_need <= 200 * 1024;
    const int iters = frames_per_col;
    int nsplit = std::max(1, (iters + 255) / 256);
    const long long target = (long long)p->sms * (in_smem ? 2 : 2);
    if ((long long)ncs * nsplit < target) nsplit = (int)std::min<long long>(std::max(1, iters / 4), (target + ncs - 1) / ncs);
    nsplit = std::max(nsplit, 1);
    const int chunk = (iters + nsplit - 1) / nsplit;
    nsplit = (iters + chunk - 1) / chunk;
    a.gpc = 1;
    a.chunk = chunk;
    a.nsplit = nsplit;
    if (nsplit > 1) {
        const size_t need = (size_t)ncs * nsplit * N;
        size_t have_b = p->partial_elems * sizeof(float);
        int rc = ensure_buffer((void**)&p->d_partial, &have_b, need * sizeof(float));
        p->partial_elems = have_b / sizeof(float);
        if (rc) return rc;
        a.partial = p->d_partial;
    }
    const long long items = (long long)ncs * nsplit;
    long long grid = std::min<long long>(items, (long long)p->sms * 4);
    float2* gwork = nullptr;
    float* gacc = nullptr;
    size_t smem = 0;
    if (in_smem) {
        smem = smem_need;
        CUDA_TRY(cudaFuncSetAttribute((const void*)sti_bluestein_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      200 * 1024));
    } else {
        grid = std::min<long long>(items, (long long)p->sms * 2);
        const size_t slab = M + (size_t)(N + 1) / 2;  // float2 units: work buffer + accumulators
        if (p->gwork_slabs < (size_t)grid * slab) {
            if (p->d_gwork) { CUDA_TRY(cudaStreamSynchronize(st)); cudaFree(p->d_gwork); cudaFree(p->d_gacc); }
            p->d_gwork = nullptr; p->d_gacc = nullptr; p->gwork_slabs = 0;
            cudaError_t e1 = cudaMalloc(&p->d_gwork, sizeof(float2) * (size_t)grid * M);
            cudaError_t e2 = cudaMalloc(&p->d_gacc, sizeof(float) * (size_t)grid * N);
            if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(PSG_ERR_NOMEM, "scratch for nfft=%d failed", N);
            p->gwork_slabs = (size_t)grid * slab;
        }
        gwork = p->d_gwork;
        gacc = p->d_gacc;
    }
    BluesteinArgs b;
    b.aw = p->d_aw;
    b.bbr = p->d_bbr;
    b.twm = p->d_twm;
    b.n = N;
    b.logm = p->logm;
    void* args[] = {(void*)&a, (void*)&b, (void*)&gwork, (void*)&gacc};
    CUDA_TRY(cudaLaunchKernel((const void*)sti_bluestein_kernel, dim3((unsigned)grid), dim3(256), args, smem, st));
    g_launches++;
    if (nsplit > 1) {
        const size_t total = (size_t)ncs * N;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)p->sms * 16);
        sti_finalize_kernel<<<blocks, 256, 0, st>>>(a.partial, nsplit, N, (size_t)ncs, a.scale, a.eps, a.out_lin, a.out_db);
        g_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return PSG_OK;
}

static int iq_bytes(int iq_type) { return iq_type == PSG_IQ_C64 ? 8 : iq_type == PSG_IQ_CI16 ? 4 : iq_type == PSG_IQ_CI8 ? 2 : 0; }

extern "C" int psg_sti_run(psg_plan* p, const void* iq_dev, int64_t sample_stride, int64_t sub_stride, int nsub,
                           const int64_t* col_offset_dev, int ncol, int frames_per_col, int64_t hop, float in_scale,
                           float eps, float* out_lin_dev, float* out_db_dev, void* cuda_stream) {
    return psg_sti_run_typed(p, iq_dev, PSG_IQ_C64, sample_stride, sub_stride, nsub, col_offset_dev, ncol, frames_per_col,
                             hop, in_scale, eps, out_lin_dev, out_db_dev, cuda_stream);
}

// column offsets -> [0, max_ok], and a flag when one had to be moved (one CTA; the table is a few thousand entries)
__global__ void __launch_bounds__(256) check_offsets_kernel(const long long* __restrict__ off, int ncol, long long max_ok,
                                                            long long* __restrict__ out, int* __restrict__ flag) {
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    int mine = 0;
    for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
        const long long o = off[c];
        const long long k = o < 0 ? 0 : (o > max_ok ? max_ok : o);
        out[c] = k;
        mine |= (k != o);
    }
    if (mine) bad = 1;
    __syncthreads();
    if (threadIdx.x == 0) *flag = bad;
}

extern "C" int psg_sti_run_checked(psg_plan* p, const void* iq_dev, int iq_type, int64_t iq_elems, int64_t sample_stride,
                                   int64_t sub_stride, int nsub, const int64_t* col_offset_dev, int ncol, int frames_per_col,
                                   int64_t hop, float in_scale, float eps, float* out_lin_dev, float* out_db_dev,
                                   int32_t* oob_flag_dev, void* cuda_stream) {
    if (!p) return fail(PSG_ERR_ARG, "psg_sti_run_checked: plan is NULL");
    if (!col_offset_dev || !oob_flag_dev) return fail(PSG_ERR_ARG, "psg_sti_run_checked: NULL pointer");
    if (nsub < 1 || ncol < 1 || frames_per_col < 1 || sample_stride < 1 || sub_stride < 0 || (frames_per_col > 1 && hop < 1))
        return fail(PSG_ERR_ARG, "psg_sti_run_checked: bad shape/stride");
    // elements one column reads past its offset (as in psg_sti_host)
    const long long col_extent = ((long long)(frames_per_col - 1) * hop + (p->nfft - 1)) * sample_stride +
                                 (long long)(nsub - 1) * sub_stride + 1;
    if (iq_elems < col_extent)
        return fail(PSG_ERR_ARG, "psg_sti_run_checked: a column needs %lld elements, the array has %lld", col_extent,
                    (long long)iq_elems);
    PSG_ON_DEVICE(p->device);
    {
        size_t have = p->offchk_elems * 8;
        const int rc = ensure_buffer((void**)&p->d_offchk, &have, (size_t)ncol * 8);
        p->offchk_elems = have / 8;
        if (rc) return rc;
    }
    check_offsets_kernel<<<1, 256, 0, (cudaStream_t)cuda_stream>>>((const long long*)col_offset_dev, ncol,
                                                                   (long long)iq_elems - col_extent, p->d_offchk, oob_flag_dev);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return psg_sti_run_typed(p, iq_dev, iq_type, sample_stride, sub_stride, nsub, (const int64_t*)p->d_offchk, ncol,
                             frames_per_col, hop, in_scale, eps, out_lin_dev, out_db_dev, cuda_stream);
}

extern "C" int psg_sti_run_typed(psg_plan* p, const void* iq_dev, int iq_type, int64_t sample_stride, int64_t sub_stride,
                                 int nsub, const int64_t* col_offset_dev, int ncol, int frames_per_col, int64_t hop,
                                 float in_scale, float eps, float* out_lin_dev, float* out_db_dev, void* cuda_stream) {
    if (!p) return fail(PSG_ERR_ARG, "psg_sti_run: plan is NULL");
    const int iqb = iq_bytes(iq_type);
    if (!iqb) return fail(PSG_ERR_ARG, "psg_sti_run: unknown iq_type %d", iq_type);
    if (!iq_dev || !col_offset_dev) return fail(PSG_ERR_ARG, "psg_sti_run: NULL input pointer");
    if (!out_lin_dev && !out_db_dev) return fail(PSG_ERR_ARG, "psg_sti_run: both outputs are NULL");
    if (nsub < 1 || ncol < 1 || frames_per_col < 1)
        return fail(PSG_ERR_ARG, "psg_sti_run: nsub=%d ncol=%d frames_per_col=%d must be >= 1", nsub, ncol, frames_per_col);
    if (sample_stride < 1) return fail(PSG_ERR_ARG, "psg_sti_run: sample_stride=%lld must be >= 1", (long long)sample_stride);
    if (frames_per_col > 1 && hop < 1) return fail(PSG_ERR_ARG, "psg_sti_run: hop=%lld must be >= 1", (long long)hop);
    if ((reinterpret_cast<uintptr_t>(iq_dev) & (uintptr_t)(iqb - 1)) != 0)
        return fail(PSG_ERR_ARG, "psg_sti_run: iq_dev must be %d-byte aligned", iqb);
    if ((long long)ncol * nsub > (1ll << 30)) return fail(PSG_ERR_ARG, "psg_sti_run: too many columns");
    if (((reinterpret_cast<uintptr_t>(out_lin_dev) | reinterpret_cast<uintptr_t>(out_db_dev)) & 15) != 0 && p->nfft % 4 == 0)
        return fail(PSG_ERR_ARG, "psg_sti_run: output images must be 16-byte aligned");
    PSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int N = p->nfft;
    const int ncs = ncol * nsub;

    StiArgs a;
    a.iq = iq_dev;
    a.iq_type = iq_type;
    a.sample_stride = sample_stride;
    a.sub_stride = sub_stride;
    a.hop_elems = hop * sample_stride;
    a.col_off = (const long long*)col_offset_dev;
    a.ncol = ncol;
    a.nsub = nsub;
    a.nfr = frames_per_col;
    a.win = p->d_win;
    a.tw = p->d_tw;
    a.twp = nullptr;
    a.scale = (float)((double)in_scale * (double)in_scale / (double)frames_per_col);
    a.eps = eps;
    a.out_lin = out_lin_dev;
    a.out_db = out_db_dev;
    a.partial = nullptr;
    a.gpc = 1;
    a.chunk = frames_per_col;
    a.nsplit = 1;

    if (p->logn < 0) return run_bluestein(p, a, ncs, frames_per_col, st);
    const bool tma_ok = sample_stride == 1 && (reinterpret_cast<uintptr_t>(iq_dev) & 15) == 0;
    const Variant* v = nullptr;
    if (!g_force_generic.load()) {
        v = pick_variant(p->logn, tma_ok, iq_type);
        bool force_split = false, force_cluster = false;
        int whole_nst = 0;   // whole-frame path forced with this many ring stages
        int whole_rows = 0;  // ... and this many rows per CTA (0: the default form of the size)
        bool whole_f = false;  // 16384: whole-frame kernel with the radix-2 passes in registers (sti_whole16.cuh)
        // Measured defaults (profiles/r01_big_nfft_cluster_vs_split.txt, 4 and 12 GB): the cluster kernel with
        // the exchange in L2 and rows loaded to registers wins at 16384 (31 % vs 26 % of the HBM peak) and
        // 32768 (29 % vs 27 %); at 65536 the 16-CTA clusters fill only 112 of the 148 SMs and the
        // three-launch split path stays ahead (27 % vs 25 %).
        int rowtma = g_cluster_rowtma.load();
        {
            std::lock_guard<NoMutex> lk(g_variant_mu);
            force_split = g_variant_override == "split";
            force_cluster = g_variant_override == "cluster" || g_variant_override == "cluster_ldg" || g_variant_override == "cluster_dsmem";
            if (g_variant_override == "cluster") rowtma = 1;
            if (g_variant_override == "cluster_ldg") rowtma = 0;
            if (g_variant_override == "cluster_dsmem") rowtma = 2;
            if (g_variant_override == "whole") whole_nst = 4;
            if (g_variant_override == "whole_r2") { whole_nst = 4; whole_rows = 2; }
            if (g_variant_override == "whole_r4") { whole_nst = 4; whole_rows = 4; }
            if (g_variant_override == "whole_s2") whole_nst = 2;
            if (g_variant_override == "whole_s8") whole_nst = 8;
            whole_f = g_variant_override == "whole_f";
        }
        bool force_r32 = false, other_path = false;
        {
            std::lock_guard<NoMutex> lk(g_variant_mu);
            force_r32 = g_variant_override == "r32";
            other_path = !g_variant_override.empty() && !force_r32;
        }
        // 16384 / 32768 / 65536 points: three-pass radix-32 kernels (sti_r32.cuh), the measured default;
        // 8192 points: the same kernel with 256 threads, two CTAs per SM (selectable: "r32")
        // (default there for one frame per column: 333 against 250 Gsamples/s; with integration the 16x16x2x16
        // kernel stays ahead, 61 % against 55 % of the HBM peak)
        const bool r32_8192 = p->logn == 13 && frames_per_col == 1 && !other_path && g_use_multi.load();
        if (((p->logn >= 14 && p->logn <= 16 && !v && !other_path) || r32_8192 || (force_r32 && p->logn >= 13 && p->logn <= 16)) && tma_ok) {
            bool ran = false;
            const int rcr = run_r32(p, a, ncs, frames_per_col, st, &ran);
            if (rcr || ran) return rcr;
        }
        if (whole_f && p->logn == 14 && tma_ok) return run_whole16(p, a, ncs, frames_per_col, st);
        // 16384 points: the whole frame in one SM (sti_whole.cuh) is the measured default (39 % of the HBM peak
        // against 30 % for the 4-CTA cluster kernel, profiles/r01_whole_frame_16384.txt)
        // 32768 points: the same kernel on clusters of two CTAs (36 % against 29 %); at 65536 (clusters of four: 24 %)
        // the split path stays ahead (27 %)
        if ((p->logn == 14 || p->logn == 15) && tma_ok && !v && !force_split && !force_cluster && !whole_nst) whole_nst = 4;
        if (p->logn >= 14 && p->logn <= 16 && tma_ok && whole_nst == 4 && (whole_rows == 2 || p->logn >= 15)) {
            bool ran = false;
            const int rcw = run_wholec(p, a, ncs, frames_per_col, whole_rows ? whole_rows : 4, st, &ran);
            if (rcw || ran) return rcw;
        }
        if ((p->logn == 13 || p->logn == 14) && tma_ok && whole_nst) return run_whole(p, a, ncs, frames_per_col, whole_nst, st);
        const bool cluster_default = false;  // superseded by the whole-frame kernels (kept selectable)
        if (p->logn >= 13 && p->logn <= 16 && tma_ok && !force_split && (force_cluster || cluster_default)) {
            // one kernel, r0 CTAs per frame (sti_cluster.cuh); contiguous, 16-byte aligned recordings
            bool ran = false;
            const int rcc = run_cluster(p, a, ncs, frames_per_col, rowtma, st, &ran);
            if (rcc || ran) return rcc;
        }
        if (p->logn >= 13 && p->logn <= 16 && (force_split || force_cluster || !v)) return run_split(p, a, ncs, frames_per_col, st);
    }

    if (v) {
        // per-pass twiddle tables: one layout per radix set; rebuilt when the variant changes
        std::string base = v->name;
        if (v->multi) base.resize(base.size() - 2);  // "_m" twins share the tables of their base variant
        if (p->twp_name != base || !p->d_twp) {
            if (p->d_twp) { CUDA_TRY(cudaStreamSynchronize(st)); cudaFree(p->d_twp); p->d_twp = nullptr; }
            int rc = upload_pass_tables(v, &p->d_twp);
            if (rc) return rc;
            p->twp_name = base;
        }
        snprintf(p->variant_name, sizeof(p->variant_name), "%s", v->name);
        a.twp = p->d_twp;
        const Variant* used = v;
        const int rcl = launch_fused(p, v, a, ncs, frames_per_col, 16, st, &used);
        snprintf(p->variant_name, sizeof(p->variant_name), "%s", used->name);
        return rcl;
    }
    {
        // generic radix-2 path
        snprintf(p->variant_name, sizeof(p->variant_name), "generic_radix2");
        const bool in_smem = N <= 16384;
        const int iters = frames_per_col;
        int nsplit = std::max(1, (iters + 255) / 256);
        const long long target = (long long)p->sms * 8;
        if ((long long)ncs * nsplit < target) nsplit = (int)std::min<long long>(std::max(1, iters / 4), (target + ncs - 1) / ncs);
        nsplit = std::max(nsplit, 1);
        int chunk = (iters + nsplit - 1) / nsplit;
        nsplit = (iters + chunk - 1) / chunk;
        a.gpc = 1;
        a.chunk = chunk;
        a.nsplit = nsplit;
        if (nsplit > 1) {
            const size_t need = (size_t)ncs * nsplit * N;
            size_t have_b = p->partial_elems * sizeof(float);
            int rc = ensure_buffer((void**)&p->d_partial, &have_b, need * sizeof(float));
            p->partial_elems = have_b / sizeof(float);
            if (rc) return rc;
            a.partial = p->d_partial;
        }
        const long long items = (long long)ncs * nsplit;
        size_t smem = 0;
        float2* gwork = nullptr;
        float* gacc = nullptr;
        long long grid = items;
        if (in_smem) {
            smem = (size_t)N * 12;
            cudaError_t e = cudaFuncSetAttribute((const void*)sti_generic_kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 12);
            if (e != cudaSuccess) return fail(PSG_ERR_CUDA, "cudaFuncSetAttribute(generic): %s", cudaGetErrorString(e));
            grid = std::min<long long>(items, (long long)p->sms * 16);
        } else {
            grid = std::min<long long>(items, (long long)p->sms * 2);
            if (p->gwork_slabs < (size_t)grid) {
                if (p->d_gwork) { CUDA_TRY(cudaStreamSynchronize(st)); cudaFree(p->d_gwork); cudaFree(p->d_gacc); }
                p->d_gwork = nullptr; p->d_gacc = nullptr; p->gwork_slabs = 0;
                cudaError_t e1 = cudaMalloc(&p->d_gwork, sizeof(float2) * (size_t)grid * N);
                cudaError_t e2 = cudaMalloc(&p->d_gacc, sizeof(float) * (size_t)grid * N);
                if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(PSG_ERR_NOMEM, "scratch for nfft=%d failed", N);
                p->gwork_slabs = (size_t)grid;
            }
            gwork = p->d_gwork;
            gacc = p->d_gacc;
        }
        int logn = p->logn;
        void* args[] = {(void*)&a, (void*)&logn, (void*)&gwork, (void*)&gacc};
        CUDA_TRY(cudaLaunchKernel((const void*)sti_generic_kernel, dim3((unsigned)grid), dim3(256), args, smem, st));
        g_launches++;
    }
    if (a.nsplit > 1) {
        const size_t total = (size_t)ncs * N;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)p->sms * 16);
        sti_finalize_kernel<<<blocks, 256, 0, st>>>(a.partial, a.nsplit, N, (size_t)ncs, a.scale, eps, out_lin_dev,
                                                    out_db_dev);
        g_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return PSG_OK;
}

extern "C" int psg_median_time(psg_plan* p, const float* img_dev, int nsub, int ncol, int nfft, float eps,
                               float* med_lin_dev, float* med_db_dev, void* cuda_stream) {
    if (!p) return fail(PSG_ERR_ARG, "psg_median_time: plan is NULL");
    if (!img_dev || (!med_lin_dev && !med_db_dev)) return fail(PSG_ERR_ARG, "psg_median_time: NULL pointer");
    if (nsub < 1 || ncol < 1 || nfft < 1) return fail(PSG_ERR_ARG, "psg_median_time: bad shape");
    PSG_ON_DEVICE(p->device);
    // Default: warp-per-bin selection (median_select_kernel).  Rows padded to a multiple of 128 keys + 4 (the
    // transposing load is then bank-conflict free).  Bins per CTA (= warps per CTA): what keeps the most warps
    // resident -- all the keys of a bin stay in shared memory, so at 3600 columns an SM holds 15 bins.
    {
        const int rowlen = ((ncol + 127) / 128) * 128 + 4;  // whole trips of 32 x 128-bit loads over a complete row
        const size_t row_bytes = (size_t)rowlen * 4, sm_total = 227 * 1024;
        int bpc = 0;
        long long best = 0;
        for (int cand_b : {8, 4, 2}) {
            const size_t smem = cand_b * row_bytes;
            if (smem > 220 * 1024) continue;
            const long long ctas = std::min<long long>(32, (long long)(sm_total / (smem + 1024)));
            // rows of two floats read a quarter of every 32-byte sector (four CTAs share it through L2): only when
            // that buys clearly more resident warps.  (3 / 5 / 6 bins per CTA -- 15 instead of 12 warps per SM at 3600
            // columns -- measured slower: their rows straddle the sectors, profiles/r01_median_kernel_experiment.txt)
            const long long warps = std::min<long long>(64, ctas * cand_b) * (cand_b == 2 ? 4 : 5);
            if (warps > best) { best = warps; bpc = cand_b; }
        }
        if (bpc && !g_force_generic.load()) {
            const size_t smem = bpc * row_bytes;
            const void* fn = bpc == 8 ? (const void*)median_select_kernel<8>
                             : bpc == 4 ? (const void*)median_select_kernel<4>
                                        : (const void*)median_select_kernel<2>;
            static thread_local const void* q_fn = nullptr;
            static thread_local int q_dev = -1;
            if (q_fn != fn || q_dev != p->device) {
                CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                q_fn = fn;
                q_dev = p->device;
            }
            const long long blocks = (long long)nsub * ((nfft + bpc - 1) / bpc);
            void* args[] = {(void*)&img_dev, (void*)&nsub, (void*)&ncol, (void*)&nfft, (void*)&rowlen, (void*)&eps,
                            (void*)&med_lin_dev, (void*)&med_db_dev};
            CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)blocks), dim3(bpc * 32), args, smem, (cudaStream_t)cuda_stream));
            g_launches++;
            CUDA_TRY(cudaGetLastError());
            return PSG_OK;
        }
    }
    // Fallback (more columns than a tile holds; also what psg_debug_set_force_generic selects, as the cross-check of
    // the selection kernel): CTA-wide bisection, tiled while [ncol][8] keys fit, else re-reading L2.
    const size_t smem_max = 222 * 1024;
    const size_t hdr = 3 * 256 * sizeof(unsigned);
    int bpc = 8;
    for (int cand = 32; cand >= 8; cand /= 2) {
        const long long ctas = (long long)nsub * ((nfft + cand - 1) / cand);
        if (hdr + (size_t)ncol * cand * 4 <= smem_max && (ctas >= p->sms || cand == 8)) { bpc = cand; break; }
    }
    const bool tiled = hdr + (size_t)ncol * bpc * 4 <= smem_max;
    const size_t smem = hdr + (tiled ? (size_t)ncol * bpc * 4 : 0);
    const void* fn = nullptr;
    if (tiled) fn = bpc == 32 ? (const void*)median_time_kernel<32, true> : bpc == 16 ? (const void*)median_time_kernel<16, true>
                                                                                      : (const void*)median_time_kernel<8, true>;
    else fn = (const void*)median_time_kernel<8, false>;
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    const long long blocks = (long long)nsub * ((nfft + bpc - 1) / bpc);
    void* args[] = {(void*)&img_dev, (void*)&nsub, (void*)&ncol, (void*)&nfft, (void*)&eps, (void*)&med_lin_dev,
                    (void*)&med_db_dev};
    CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)blocks), dim3(256), args, smem, (cudaStream_t)cuda_stream));
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return PSG_OK;
}

extern "C" int psg_minmax_time(psg_plan* p, const float* img_dev, int nsub, int ncol, int nfft, float eps,
                               float* min_lin_dev, float* max_lin_dev, float* min_db_dev, float* max_db_dev,
                               void* cuda_stream) {
    if (!p) return fail(PSG_ERR_ARG, "psg_minmax_time: plan is NULL");
    if (!img_dev || (!min_lin_dev && !max_lin_dev && !min_db_dev && !max_db_dev))
        return fail(PSG_ERR_ARG, "psg_minmax_time: NULL pointer");
    if (nsub < 1 || ncol < 1 || nfft < 1) return fail(PSG_ERR_ARG, "psg_minmax_time: bad shape");
    PSG_ON_DEVICE(p->device);
    const long long blocks = (long long)nsub * ((nfft + 31) / 32);
    minmax_time_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream>>>(img_dev, nsub, ncol, nfft, eps, min_lin_dev,
                                                                               max_lin_dev, min_db_dev, max_db_dev);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return PSG_OK;
}

extern "C" int psg_gather_bins(psg_plan* p, const float* img_dev, int64_t rows, int nfft, const int32_t* idx_dev, int count,
                               float clamp_lo, float clamp_hi, float* out_dev, void* cuda_stream) {
    if (!p) return fail(PSG_ERR_ARG, "psg_gather_bins: plan is NULL");
    if (!img_dev || !idx_dev || !out_dev) return fail(PSG_ERR_ARG, "psg_gather_bins: NULL pointer");
    if (rows < 1 || nfft < 1 || count < 1) return fail(PSG_ERR_ARG, "psg_gather_bins: bad shape");
    PSG_ON_DEVICE(p->device);
    const size_t total = (size_t)rows * (size_t)count;
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)p->sms * 16);
    gather_bins_kernel<<<blocks, 256, 0, (cudaStream_t)cuda_stream>>>(img_dev, (size_t)rows, nfft, idx_dev, count, clamp_lo,
                                                                     clamp_hi, out_dev);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return PSG_OK;
}

extern "C" int psg_sti_host(psg_plan* p, const void* iq_host, int64_t iq_host_elems, int64_t sample_stride,
                            int64_t sub_stride, int nsub, const int64_t* col_offset_host, int ncol,
                            int frames_per_col, int64_t hop, float in_scale, float eps, float* out_lin_host,
                            float* out_db_host, float* med_lin_host, float* med_db_host) {
    return psg_sti_host_typed(p, iq_host, PSG_IQ_C64, iq_host_elems, sample_stride, sub_stride, nsub, col_offset_host,
                              ncol, frames_per_col, hop, in_scale, eps, out_lin_host, out_db_host, med_lin_host,
                              med_db_host);
}

extern "C" int psg_sti_host_typed(psg_plan* p, const void* iq_host, int iq_type, int64_t iq_host_elems,
                                  int64_t sample_stride, int64_t sub_stride, int nsub, const int64_t* col_offset_host,
                                  int ncol, int frames_per_col, int64_t hop, float in_scale, float eps,
                                  float* out_lin_host, float* out_db_host, float* med_lin_host, float* med_db_host) {
    if (!p) return fail(PSG_ERR_ARG, "psg_sti_host: plan is NULL");
    const int iqb = iq_bytes(iq_type);
    if (!iqb) return fail(PSG_ERR_ARG, "psg_sti_host: unknown iq_type %d", iq_type);
    if (!iq_host || !col_offset_host) return fail(PSG_ERR_ARG, "psg_sti_host: NULL input pointer");
    if (!out_lin_host && !out_db_host && !med_lin_host && !med_db_host)
        return fail(PSG_ERR_ARG, "psg_sti_host: no output requested");
    if (nsub < 1 || ncol < 1 || frames_per_col < 1 || sample_stride < 1 || sub_stride < 0)
        return fail(PSG_ERR_ARG, "psg_sti_host: bad shape/stride");
    const int N = p->nfft;
    // extent of one column's reads (elements past its offset), and the span all columns touch
    const long long col_extent = ((long long)(frames_per_col - 1) * hop + (N - 1)) * sample_stride +
                                 (long long)(nsub - 1) * sub_stride + 1;
    long long lo = col_offset_host[0], hi = col_offset_host[0];
    for (int c = 1; c < ncol; ++c) {
        lo = std::min<long long>(lo, col_offset_host[c]);
        hi = std::max<long long>(hi, col_offset_host[c]);
    }
    if (lo < 0 || hi + col_extent > iq_host_elems)
        return fail(PSG_ERR_ARG, "psg_sti_host: columns reach [%lld, %lld) outside the %lld-element host array", lo,
                    hi + col_extent, (long long)iq_host_elems);
    PSG_ON_DEVICE(p->device);
    cudaStream_t st = p->stream;
    // keep the device copy 16-byte aligned relative to the host element parity so that aligned
    // host frames stay aligned on the device
    const long long al_mask = ~(long long)(16 / iqb - 1);
    const long long lo_al = lo & al_mask;
    const size_t span = (size_t)(hi + col_extent - lo_al);
    // Recordings larger than the chunk size are streamed in groups of consecutive columns whose
    // touched span fits a chunk -- when that actually shrinks the copy (column spans that overlap
    // almost entirely, e.g. the reference's (rows, ntime, nsub) array, are copied once instead).
    struct HostChunk { int c0, c1; long long lo_al, span; };
    std::vector<HostChunk> chunks;
    const long long cap = std::max<long long>(1, g_host_chunk_bytes.load() / iqb);
    if ((long long)span > cap && ncol > 1) {
        long long sum = 0, biggest = 0;
        for (int c0 = 0; c0 < ncol;) {
            long long clo = col_offset_host[c0], chi = clo;
            int c1 = c0 + 1;
            while (c1 < ncol) {
                const long long nlo = std::min<long long>(clo, col_offset_host[c1]), nhi = std::max<long long>(chi, col_offset_host[c1]);
                if (nhi + col_extent - (nlo & al_mask) > cap) break;
                clo = nlo;
                chi = nhi;
                ++c1;
            }
            const long long cl = clo & al_mask;
            chunks.push_back({c0, c1, cl, chi + col_extent - cl});
            sum += chunks.back().span;
            biggest = std::max(biggest, chunks.back().span);
            c0 = c1;
        }
        if (chunks.size() < 2 || sum > (long long)span + (long long)span / 4) chunks.clear();  // no gain: copy once
        if (!chunks.empty()) {
            int rcb = ensure_buffer(&p->d_in, &p->in_bytes, (size_t)biggest * iqb + 32);
            if (rcb) return rcb;
            rcb = ensure_buffer(&p->d_in2, &p->in2_bytes, (size_t)biggest * iqb + 32);
            if (rcb) return rcb;
        }
    }
    int rc = PSG_OK;
    if (chunks.empty()) {
        rc = ensure_buffer(&p->d_in, &p->in_bytes, span * iqb + 32);
        if (rc) return rc;
    }
    {
        size_t have = p->off_elems * 8;
        rc = ensure_buffer((void**)&p->d_off, &have, (size_t)ncol * 8);
        p->off_elems = have / 8;
        if (rc) return rc;
    }
    std::vector<long long> rel(ncol);
    for (int c = 0; c < ncol; ++c) rel[c] = col_offset_host[c] - lo_al;
    const size_t img_elems = (size_t)nsub * ncol * N, med_elems = (size_t)nsub * N;
    const bool need_lin = out_lin_host || med_lin_host || med_db_host;
    const size_t want[4] = {need_lin ? img_elems : 0, out_db_host ? img_elems : 0, med_lin_host ? med_elems : 0,
                            med_db_host ? med_elems : 0};
    for (int i = 0; i < 4; ++i) {
        if (!want[i]) continue;
        size_t have = p->out_elems[i] * 4;
        rc = ensure_buffer((void**)&p->d_out[i], &have, want[i] * 4);
        p->out_elems[i] = have / 4;
        if (rc) return rc;
    }
    if (chunks.size() <= 1) {
        CUDA_TRY(cudaMemcpyAsync(p->d_off, rel.data(), (size_t)ncol * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(p->d_in, (const char*)iq_host + (size_t)lo_al * iqb, span * iqb, cudaMemcpyHostToDevice, st));
        rc = psg_sti_run_typed(p, p->d_in, iq_type, sample_stride, sub_stride, nsub, (const int64_t*)p->d_off, ncol,
                               frames_per_col, hop, in_scale, eps, need_lin ? p->d_out[0] : nullptr,
                               out_db_host ? p->d_out[1] : nullptr, st);
        if (rc) return rc;
    } else {
        // streamed: chunk j+1 crosses PCIe on the copy stream while the kernels of chunk j run on the
        // plan's stream (all kernels stay on one stream: they share the plan's scratch); two staging
        // buffers, a "copied" and a "free" event each.  A chunk's columns land in a contiguous slab of
        // the image per sub-channel, so sub-channels run as separate launches.
        if (!p->copy_stream) {
            CUDA_TRY(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) {
                CUDA_TRY(cudaEventCreateWithFlags(&p->ev_copied[i], cudaEventDisableTiming));
                CUDA_TRY(cudaEventCreateWithFlags(&p->ev_free[i], cudaEventDisableTiming));
            }
        }
        for (const HostChunk& ch : chunks)
            for (int c = ch.c0; c < ch.c1; ++c) rel[c] = col_offset_host[c] - ch.lo_al;
        CUDA_TRY(cudaMemcpyAsync(p->d_off, rel.data(), (size_t)ncol * 8, cudaMemcpyHostToDevice, st));
        void* bufs[2] = {p->d_in, p->d_in2};
        for (size_t j = 0; j < chunks.size(); ++j) {
            const HostChunk& ch = chunks[j];
            const int b = (int)(j & 1);
            if (j >= 2) CUDA_TRY(cudaStreamWaitEvent(p->copy_stream, p->ev_free[b], 0));
            CUDA_TRY(cudaMemcpyAsync(bufs[b], (const char*)iq_host + (size_t)ch.lo_al * iqb, (size_t)ch.span * iqb,
                                     cudaMemcpyHostToDevice, p->copy_stream));
            CUDA_TRY(cudaEventRecord(p->ev_copied[b], p->copy_stream));
            CUDA_TRY(cudaStreamWaitEvent(st, p->ev_copied[b], 0));
            const int nc = ch.c1 - ch.c0;
            for (int sub = 0; sub < nsub; ++sub) {
                const size_t o = ((size_t)sub * ncol + ch.c0) * N;
                rc = psg_sti_run_typed(p, (const char*)bufs[b] + (size_t)sub * sub_stride * iqb, iq_type, sample_stride, 0, 1,
                                       (const int64_t*)p->d_off + ch.c0, nc, frames_per_col, hop, in_scale, eps,
                                       need_lin ? p->d_out[0] + o : nullptr, out_db_host ? p->d_out[1] + o : nullptr, st);
                if (rc) {
                    cudaStreamSynchronize(p->copy_stream);
                    cudaStreamSynchronize(st);
                    return rc;
                }
            }
            CUDA_TRY(cudaEventRecord(p->ev_free[b], st));
        }
    }
    if (med_lin_host || med_db_host) {
        rc = psg_median_time(p, p->d_out[0], nsub, ncol, N, eps, med_lin_host ? p->d_out[2] : nullptr,
                             med_db_host ? p->d_out[3] : nullptr, st);
        if (rc) return rc;
    }
    if (out_lin_host) CUDA_TRY(cudaMemcpyAsync(out_lin_host, p->d_out[0], img_elems * 4, cudaMemcpyDeviceToHost, st));
    if (out_db_host) CUDA_TRY(cudaMemcpyAsync(out_db_host, p->d_out[1], img_elems * 4, cudaMemcpyDeviceToHost, st));
    if (med_lin_host) CUDA_TRY(cudaMemcpyAsync(med_lin_host, p->d_out[2], med_elems * 4, cudaMemcpyDeviceToHost, st));
    if (med_db_host) CUDA_TRY(cudaMemcpyAsync(med_db_host, p->d_out[3], med_elems * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (p->copy_stream) CUDA_TRY(cudaStreamSynchronize(p->copy_stream));
    return PSG_OK;
}
