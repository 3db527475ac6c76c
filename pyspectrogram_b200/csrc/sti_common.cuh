// Primitives shared by every kernel family: argument block, sample decoding, dB conversion, mbarrier /
// bulk-copy (TMA) / cluster / distributed-shared-memory wrappers.  No __global__ definitions here, so any
// translation unit may include it.
#pragma once
#include <stdint.h>
#include "cplx.cuh"

struct StiArgs {
    const void* iq;   // complex samples: fp32 pairs (c64), int16 pairs (ci16) or int8 pairs (ci8)
    int iq_type;      // PSG_IQ_* (runtime switch for the generic / split kernels; tuned kernels are templated)
    long long sample_stride;  // elements between consecutive samples
    long long sub_stride;     // elements between sub-channels
    long long hop_elems;      // hop * sample_stride
    const long long* col_off; // [ncol] element offset of each column's first sample
    int ncol, nsub;
    int nfr;     // frames per column
    int chunk;   // frames per work item (per column)
    int nsplit;  // work items per column = ceil(nfr / chunk)
    int gpc;     // frame groups cooperating on one column (power of two, divides F)
    const float* win;   // [N]  w[n] / sum(w)
    const float2* tw;   // [N]  exp(-2*pi*j*m/N)   (generic kernels)
    const float2* twp;  // per-pass tables, concatenated in pass order (tuned kernels)
    float scale;        // in_scale^2 / nfr
    float eps;
    float* out_lin;     // [nsub][ncol][N] or null
    float* out_db;      // [nsub][ncol][N] or null
    float* partial;     // [nsub*ncol][nsplit][N] raw sums when nsplit > 1
    int cb;             // MULTI kernels: consecutive column blocks per CTA
};

enum { PSG_LOADER_LDG = 0, PSG_LOADER_TMA = 1 };

PSG_DEV float2 ldg_stream(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

// Raw integer IQ (Digital RF's native complex int16 / int8): element size
// and decoding.  The full-scale factor 1/ref (drfProc.py:129, get_ref :182-201) is not applied per
// sample: it is folded into the power scale in_scale^2 of the epilogue.
enum { IQ_C64 = 0, IQ_CI16 = 1, IQ_CI8 = 2 };
template <int IQT> struct IqBytes { static constexpr int value = (IQT == IQ_C64) ? 8 : (IQT == IQ_CI16) ? 4 : 2; };
PSG_DEV float2 decode_ci16(uint32_t r) {
    return make_float2((float)(short)(r & 0xffffu), (float)(short)(r >> 16));
}
PSG_DEV float2 decode_ci8(uint32_t r) {
    return make_float2((float)(signed char)(r & 0xffu), (float)(signed char)((r >> 8) & 0xffu));
}
template <int IQT>
PSG_DEV float2 ldg_iq(const void* base, long long elem) {
    if constexpr (IQT == IQ_C64) {
        return ldg_stream(reinterpret_cast<const float2*>(base) + elem);
    } else if constexpr (IQT == IQ_CI16) {
        uint32_t r;
        asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(reinterpret_cast<const uint32_t*>(base) + elem));
        return decode_ci16(r);
    } else {
        unsigned short r;
        asm volatile("ld.global.nc.L1::no_allocate.b16 %0, [%1];" : "=h"(r) : "l"(reinterpret_cast<const unsigned short*>(base) + elem));
        return decode_ci8(r);
    }
}
PSG_DEV float2 ldg_iq_rt(int iqt, const void* base, long long elem) {
    if (iqt == IQ_CI16) return ldg_iq<IQ_CI16>(base, elem);
    if (iqt == IQ_CI8) return ldg_iq<IQ_CI8>(base, elem);
    return ldg_iq<IQ_C64>(base, elem);
}
// element idx of a shared-memory stage holding raw samples
template <int IQT>
PSG_DEV float2 lds_iq(const unsigned char* stage, int idx) {
    if constexpr (IQT == IQ_C64) return reinterpret_cast<const float2*>(stage)[idx];
    else if constexpr (IQT == IQ_CI16) return decode_ci16(reinterpret_cast<const uint32_t*>(stage)[idx]);
    else return decode_ci8(reinterpret_cast<const unsigned short*>(stage)[idx]);
}

// 10*log10(p + eps) (drfProc.py:308-310) as 10*log10(2) * lg2.approx(p + eps): one MUFU instead of the
// ~20-instruction log10f.  lg2.approx is within 2 ulp (2^-22 absolute near 1), i.e. <= 3e-5 dB over the
// range this path produces (>= -150 dB), against the 1e-3 dB parity bar; p + eps >= 1e-15 is never denormal.
PSG_DEV float power_to_db(float p, float eps) { return 3.0102999566398120f * __log2f(p + eps); }

__host__ __device__ constexpr int psg_pad(int pos) { return pos + 2 * (pos >> 4); }

// ---- mbarrier / bulk-copy (TMA) primitives ------------------------------------------------------
PSG_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
PSG_DEV void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
PSG_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
PSG_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP).
PSG_DEV void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// number of power-of-two exponents below R: the mid-pass twiddles W^k, k = 1..R-1, of a TWP plan are
// rebuilt in registers from W^1, W^2, W^4, W^8 (k = hb + k' -> W^hb * W^k', one complex multiply
// each) instead of being loaded: the loads cost as many LSU wavefronts as a data exchange of the
// pass, the multiplies run on the less loaded FMA pipe.
__host__ __device__ constexpr int psg_npow(int r) { return r >= 16 ? 4 : r >= 8 ? 3 : r >= 4 ? 2 : r >= 2 ? 1 : 0; }


// ---- clusters, distributed shared memory ----------------------------------------------------------
PSG_DEV unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
PSG_DEV void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
PSG_DEV void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
PSG_DEV uint32_t map_cluster(uint32_t local_addr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
PSG_DEV void st_async_cf(uint32_t raddr, cf v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(raddr),
                 "f"(v.x), "f"(v.y), "r"(rbar)
                 : "memory");
}
// a wait that cannot hang the device: a protocol error traps instead of spinning forever
PSG_DEV void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (int spins = 0;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spins > (1 << 22)) asm volatile("trap;");
    }
}

// one probe of the barrier (no spin): issued early, its latency hides behind independent arithmetic
PSG_DEV bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

// Reader count of a stage.  Relaxed on purpose: an acq_rel atomic compiles to MEMBAR.ALL.CTA, which makes
// the lane wait for every load it has in flight -- including the window loads issued ~700 clk ahead.
// Ordering comes from the data flow: a warp's reads of the stage have been consumed by its butterflies
// (and __syncwarp() has gathered the lanes) before lane 0 counts the warp, so the bulk copy the last
// counter issues cannot overtake a read.
PSG_DEV unsigned count_reader(unsigned* p) {
    unsigned old;
    asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(p)) : "memory");
    return old;
}
PSG_DEV void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
