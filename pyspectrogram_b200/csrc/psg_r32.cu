// Host side of the radix-32 whole-frame kernels (sti_r32.cuh): instantiation, occupancy query, launch.
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
template <int CL>
static const void* r32_fn(int iqt, int opt) {
    if (iqt == IQ_CI16) return (const void*)sti_r32_kernel<CL, IQ_CI16, R32_DEFAULT_OPT>;
    if (iqt == IQ_CI8) return (const void*)sti_r32_kernel<CL, IQ_CI8, R32_DEFAULT_OPT>;
    switch (opt) {
        case 1: return (const void*)sti_r32_kernel<CL, IQ_C64, 1>;
        case 2: return (const void*)sti_r32_kernel<CL, IQ_C64, 2>;
        case 3: return (const void*)sti_r32_kernel<CL, IQ_C64, 3>;
        case 4: return (const void*)sti_r32_kernel<CL, IQ_C64, 4>;
        case 5: return (const void*)sti_r32_kernel<CL, IQ_C64, 5>;
        default: return (const void*)sti_r32_kernel<CL, IQ_C64, 0>;
    }
}
template <int CL>
static size_t r32_smem(int iqt) {
    return iqt == IQ_CI16 ? R32Cfg<CL, IQ_CI16>::smem_bytes : iqt == IQ_CI8 ? R32Cfg<CL, IQ_CI8>::smem_bytes : R32Cfg<CL, IQ_C64>::smem_bytes;
}
static bool r32_pick(int logn, int iqt, const void** fn, size_t* smem, int* cl) {
    const int opt = r32_opt();
    switch (logn) {
        case 14: *fn = r32_fn<1>(iqt, opt); *smem = r32_smem<1>(iqt); *cl = 1; return true;
        case 15: *fn = r32_fn<2>(iqt, opt); *smem = r32_smem<2>(iqt); *cl = 2; return true;
        case 16: *fn = r32_fn<4>(iqt, opt); *smem = r32_smem<4>(iqt); *cl = 4; return true;
        default: return false;
    }
}
static void r32_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int cl, size_t smem, cudaStream_t st) {
    memset(cfg, 0, sizeof(*cfg));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg->blockDim = dim3(512);
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
    const void* fn;
    size_t smem;
    int cl;
    *ngroups = 0;
    if (!r32_pick(logn, iq_type, &fn, &smem, &cl)) return (int)cudaErrorInvalidValue;
    (void)device;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (cl == 1) {
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 512, smem);
        if (e != cudaSuccess) return (int)e;
        *ngroups = occ * sms;
        return (int)cudaSuccess;
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    r32_config(&cfg, attr, cl, smem, nullptr);
    cfg.gridDim = dim3((unsigned)(cl * sms));
    int nmax = 0;
    e = cudaOccupancyMaxActiveClusters(&nmax, fn, &cfg);
    if (e != cudaSuccess) {
        cudaGetLastError();
        nmax = 0;
    }
    *ngroups = nmax;
    return (int)cudaSuccess;
}

int psg_r32_launch(int logn, int iq_type, const StiArgs& a, int nitems, int ngroups, cudaStream_t st) {
    const void* fn;
    size_t smem;
    int cl;
    if (!r32_pick(logn, iq_type, &fn, &smem, &cl)) return (int)cudaErrorInvalidValue;
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
    r32_config(&cfg, attr, cl, smem, st);
    cfg.gridDim = dim3((unsigned)(ngroups * cl));
    void* args[] = {(void*)&ra};
    return (int)cudaLaunchKernelExC(&cfg, fn, args);
}
