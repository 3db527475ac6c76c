// Host side of the radix-32 whole-frame kernels (sti_r32.cuh): instantiation, occupancy query, launch.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "psg_r32.h"
#include "sti_r32.cuh"

// tuning switches of the kernel (R32_ORDER | R32_ACC2); PSG_R32_OPT in the environment overrides the default for
// A/B measurements (complex64 only: the integer twins are instantiated with the default)
static constexpr int R32_DEFAULT_OPT = R32_ORDER;
static int r32_opt() {
    static const int opt = [] {
        const char* e = getenv("PSG_R32_OPT");
        return e ? (atoi(e) & 7) : R32_DEFAULT_OPT;
    }();
    return opt;
}
template <int LOGN, int T>
static const void* r32_fn(int iqt, int opt) {
    if (iqt == IQ_CI16) return (const void*)sti_r32_kernel<LOGN, T, IQ_CI16, R32_DEFAULT_OPT>;
    if (iqt == IQ_CI8) return (const void*)sti_r32_kernel<LOGN, T, IQ_CI8, R32_DEFAULT_OPT>;
    switch (opt) {
        case 0: return (const void*)sti_r32_kernel<LOGN, T, IQ_C64, 0>;
        case 4: return (const void*)sti_r32_kernel<LOGN, T, IQ_C64, 4>;
        case 5: return (const void*)sti_r32_kernel<LOGN, T, IQ_C64, 5>;
        default: return (const void*)sti_r32_kernel<LOGN, T, IQ_C64, 1>;
    }
}
template <int LOGN, int T>
static size_t r32_smem(int iqt) {
    return iqt == IQ_CI16 ? R32Cfg<LOGN, T, IQ_CI16>::smem_bytes
           : iqt == IQ_CI8 ? R32Cfg<LOGN, T, IQ_CI8>::smem_bytes
                           : R32Cfg<LOGN, T, IQ_C64>::smem_bytes;
}
// threads per CTA: 8192 points run two CTAs of 256 threads per SM; the 256-thread cluster forms of 16384 / 32768
// (pairs / clusters of four) measured 37.6 % / 31.1 % of the HBM peak against 55.9 % / 45.1 % for 512 threads --
// twice the share of the frame crosses the SM-to-SM network -- and are not instantiated
static int r32_threads(int logn) { return logn == 13 ? 256 : 512; }
struct R32Pick {
    const void* fn;
    size_t smem;
    int cl, threads;
};
static bool r32_pick(int logn, int iqt, R32Pick* p) {
    const int opt = r32_opt(), T = r32_threads(logn);
    p->threads = T;
#define R32_CASE(LOGN, TT)                       \
    if (logn == LOGN && T == TT) {               \
        p->fn = r32_fn<LOGN, TT>(iqt, opt);      \
        p->smem = r32_smem<LOGN, TT>(iqt);       \
        p->cl = R32Geo<LOGN, TT>::CL;            \
        return true;                             \
    }
    R32_CASE(13, 256)
    R32_CASE(14, 512)
    R32_CASE(15, 512)
    R32_CASE(16, 512)
#undef R32_CASE
    return false;
}
static void r32_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int cl, int threads, size_t smem, cudaStream_t st) {
    memset(cfg, 0, sizeof(*cfg));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg->blockDim = dim3((unsigned)threads);
    cfg->dynamicSmemBytes = smem;
    cfg->stream = st;
    cfg->attrs = attr;
    cfg->numAttrs = 1;
}

static long long* g_trace = nullptr;
// copies the phase-boundary clocks of the last traced launch to the host (PSG_R32_OPT bit 2); returns the count
extern "C" int psg_r32_trace_dump(long long* out, int max_count) {
    const int n = R32_TRACE_FRAMES * 16 * R32_TRACE_EVENTS;
    if (!g_trace || max_count < n) return 0;
    cudaDeviceSynchronize();
    cudaMemcpy(out, g_trace, sizeof(long long) * n, cudaMemcpyDeviceToHost);
    return n;
}

int psg_r32_max_groups(int logn, int iq_type, int device, int sms, int* ngroups) {
    R32Pick p;
    *ngroups = 0;
    if (!r32_pick(logn, iq_type, &p)) return (int)cudaErrorInvalidValue;
    (void)device;
    cudaError_t e = cudaFuncSetAttribute(p.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return (int)e;
    if (p.cl == 1) {
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p.fn, p.threads, p.smem);
        if (e != cudaSuccess) return (int)e;
        // The occupancy API answers 1 for every kernel that allocates tensor memory, whatever its other resources
        // (measured: 1 at 32 threads and no shared memory).  Two 256-thread CTAs do share an SM -- 2 x 256 TMEM
        // columns, 2 x 112 KB, 2 x 32 K registers -- and a grid of 2 CTAs per SM measures 55 % against 51 % of the
        // HBM peak in Mode A and 61 % against 47 % in Mode R at 8192 points.
        if (p.threads == 256) occ = 2;
        *ngroups = occ * sms;
        return (int)cudaSuccess;
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    r32_config(&cfg, attr, p.cl, p.threads, p.smem, nullptr);
    cfg.gridDim = dim3((unsigned)(p.cl * sms * (512 / p.threads)));
    int nmax = 0;
    e = cudaOccupancyMaxActiveClusters(&nmax, p.fn, &cfg);
    if (e != cudaSuccess) {
        cudaGetLastError();
        nmax = 0;
    }
    *ngroups = nmax;
    return (int)cudaSuccess;
}

int psg_r32_launch(int logn, int iq_type, const StiArgs& a, int nitems, int ngroups, cudaStream_t st) {
    R32Pick p;
    if (!r32_pick(logn, iq_type, &p)) return (int)cudaErrorInvalidValue;
    R32Args ra;
    ra.s = a;
    ra.nitems = nitems;
    ra.ngroups = ngroups;
    ra.trace = nullptr;
    if (r32_opt() & R32_TRACE) {
        // debugging aid: one buffer per process, dumped by psg_r32_trace_dump()
        static long long* d_trace = nullptr;
        if (!d_trace) cudaMalloc(&d_trace, sizeof(long long) * R32_TRACE_FRAMES * 16 * R32_TRACE_EVENTS);
        ra.trace = d_trace;
        g_trace = d_trace;
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    r32_config(&cfg, attr, p.cl, p.threads, p.smem, st);
    cfg.gridDim = dim3((unsigned)(ngroups * p.cl));
    void* args[] = {(void*)&ra};
    return (int)cudaLaunchKernelExC(&cfg, p.fn, args);
}

// geometry of the kernel the launcher would pick (for the variant name)
void psg_r32_describe(int logn, int* threads, int* cl) {
    R32Pick p;
    *threads = 0;
    *cl = 0;
    if (r32_pick(logn, IQ_C64, &p)) {
        *threads = p.threads;
        *cl = p.cl;
    }
}
