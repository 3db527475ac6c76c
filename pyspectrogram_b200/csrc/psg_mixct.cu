// Host side of the compile-time mixed-radix kernels (sti_mixct.cuh): instantiation, occupancy query, launch.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "psg_mixct.h"
#include "sti_mixct.cuh"

struct MixctPick {
    const void* fn;
    int threads, groups;
    size_t smem;
    const char* name;
};

template <class PL, int F, int MINB>
static const void* mixct_fn(int iqt) {
    if (iqt == IQ_CI16) return (const void*)sti_mixct_kernel<PL, F, IQ_CI16, MINB>;
    if (iqt == IQ_CI8) return (const void*)sti_mixct_kernel<PL, F, IQ_CI8, MINB>;
    return (const void*)sti_mixct_kernel<PL, F, IQ_C64, MINB>;
}
// resident CTAs the one-group form (single-frame columns) is compiled for: as many as 128 registers per thread allow
constexpr int mixct_minb1(int t) { return 512 / t < 1 ? 1 : 512 / t; }
// PSG_MIXCT_ALT in the environment selects the alternative forms of a plan (MIXCT_ALT entries of mixct_plans.inc,
// complex64 only) for A/B measurements
static int mixct_alt() {
    static const int alt = [] {
        const char* e = getenv("PSG_MIXCT_ALT");
        return e ? atoi(e) : 0;
    }();
    return alt;
}

static bool mixct_pick(int n, int iqt, int frames_per_col, MixctPick* p) {
#define MIXCT_STR2(x) #x
#define MIXCT_STR(x) MIXCT_STR2(x)
#define MIXCT_ALT(ID, N, R0, R1, R2, R3, T, PQ, PA, TW, FD, MINB, PQ2, PA2)                                           \
    if (n == N && mixct_alt() == ID && iqt == IQ_C64) {                                                          \
        using PL = MixPlan<N, R0, R1, R2, R3, T, PQ, PA, TW, PQ2, PA2>;                                                 \
        p->fn = (const void*)sti_mixct_kernel<PL, FD, IQ_C64, MINB>;                                             \
        p->groups = FD;                                                                                          \
        p->threads = FD * T;                                                                                     \
        p->smem = (size_t)FD * PL::BUF * sizeof(float2);                                                         \
        p->name = "mixct" MIXCT_STR(N) "_alt" MIXCT_STR(ID);                                                     \
        return true;                                                                                             \
    }
#define MIXCT_PLAN(N, R0, R1, R2, R3, T, PQ, PA, TW, FD, MINB, PQ2, PA2)                                          \
    if (n == N) {                                                                                                \
        using PL = MixPlan<N, R0, R1, R2, R3, T, PQ, PA, TW, PQ2, PA2>;                                                 \
        const bool one = FD == 1 || frames_per_col < FD;                                                         \
        p->fn = one ? mixct_fn<PL, 1, (FD == 1 ? MINB : mixct_minb1(T))>(iqt) : mixct_fn<PL, FD, MINB>(iqt);     \
        p->groups = one ? 1 : FD;                                                                                \
        p->threads = p->groups * T;                                                                              \
        p->smem = (size_t)p->groups * PL::BUF * sizeof(float2);                                                  \
        p->name = (R3 > 1)   ? "mixct" MIXCT_STR(N) "_" MIXCT_STR(R0) "x" MIXCT_STR(R1) "x" MIXCT_STR(R2) "x" MIXCT_STR(R3) \
                  : (R2 > 1) ? "mixct" MIXCT_STR(N) "_" MIXCT_STR(R0) "x" MIXCT_STR(R1) "x" MIXCT_STR(R2)        \
                             : "mixct" MIXCT_STR(N) "_" MIXCT_STR(R0) "x" MIXCT_STR(R1);                         \
        return true;                                                                                             \
    }
#include "mixct_plans.inc"
#undef MIXCT_PLAN
#undef MIXCT_ALT
    return false;
}

int psg_mixct_query(int n, int iq_type, int frames_per_col, MixctInfo* info) {
    MixctPick p;
    if (!mixct_pick(n, iq_type, frames_per_col, &p)) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(p.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return (int)e;
    // (the shared-memory carve-out stays at the driver's choice: forcing the maximum takes L1 away from the twiddle
    // table, the window and the spill slots and measured 15-25 % slower on every plan)
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p.fn, p.threads, p.smem);
    if (e != cudaSuccess) return (int)e;
    info->threads = p.threads;
    info->groups = p.groups;
    info->smem = p.smem;
    info->occ = occ;
    snprintf(info->name, sizeof(info->name), "%s_f%d", p.name, p.groups);
    return (int)cudaSuccess;
}

int psg_mixct_launch(int n, int iq_type, int frames_per_col, const StiArgs& a, long long grid, cudaStream_t st) {
    MixctPick p;
    if (!mixct_pick(n, iq_type, frames_per_col, &p)) return (int)cudaErrorInvalidValue;
    void* args[] = {(void*)&a};
    return (int)cudaLaunchKernel(p.fn, dim3((unsigned)grid), dim3((unsigned)p.threads), args, p.smem, st);
}
