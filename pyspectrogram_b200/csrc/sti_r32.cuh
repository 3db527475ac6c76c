// nfft = 8192 / 16384 / 32768 / 65536 in THREE shared-memory passes: 32 points per thread, frame in place.
//
// The four-pass whole-frame kernels (sti_whole.cuh) spend eight shared-memory accesses per sample and
// stop at 36-40 % of the HBM peak with the LSU data pipe 64 % busy (profiles/r01_whole_frame_16384.txt).
// This family needs 5.5:
//   plan      N = 32 x L, L the row length; a CTA of T threads owns T columns, CL = L / T CTAs (one cluster)
//             share a frame (geometries: R32Geo in r32_math.cuh).
//             pass 0  radix 32 over n0 (elements n' + n0 L), thread <-> n' = T c + t of CTA c
//             pass 1  radix R1 = 32 (16 for L = 256) inside every row, stride S1 = L / R1
//             pass 2  S1-point DFTs on consecutive positions: radix 16 (S1 = 16), radix 32 (S1 = 32),
//                     radix 32 at stride 2 plus a radix-2 butterfly between lane pairs (SHFL) for S1 = 64
//   in place  every pass overwrites its own inputs; the buffer M (one CTA's rows, 128 KB) is not padded
//             but XOR-swizzled -- the 16-byte chunk index inside a 128-byte line is XORed with three bits
//             of the line index chosen so that the one pass whose lanes stride over lines (the last) is
//             conflict-free with LDS.128 (LDS.64 for the stride-2 form); the 64-bit accesses of passes 0
//             and 1 always cover whole lines per half-warp.
//   loader    the CTA's 32 segments of 512 samples: 24 arrive by bulk copies (UBLKCP) in a staging area
//             S (96 KB, one frame ahead, refilled by whichever warp reads it last), the other 8 straight
//             into registers with LDG issued a pass ahead (their lines are pulled into L2 when the bulk
//             copies are issued).  Integer IQ fits S whole.  T = 512: 128 KB + 96 KB fill the SM, one CTA
//             per SM; T = 256 (8192 points): 64 KB + 48 KB, two independent CTAs per SM whose phases
//             interleave (one computes while the other moves data).
//   TMEM      the tensor memory (256 KB, unused by a kernel without MMAs) is this kernel's second
//             register file: per thread 32 |X|^2 accumulators, its 32 window values and the twiddle
//             powers W^{1,2,4,8,16} of passes 0 and 1 live there (tcgen05.st once, tcgen05.ld per frame:
//             ~19 clk, no LSU wavefronts -- tools/ubench/tmem_probe.cu), so 32 complex points fit the
//             128 registers a 512-thread CTA leaves per thread.
//   cluster   CL > 1: CTA c owns rows k0 in [c 32/CL, (c+1) 32/CL); pass-0 outputs go to the owners with
//             st.async (DSMEM, counted in bytes on the owner's mbarrier); passes 1 and 2 are local.
//   sync per frame: mbarrier "S full", split "M free" (mbarrier of 16 warp arrivals, or the hardware
//             cluster barrier: arrive after the last pass, wait before the first store of the next
//             frame), one CTA barrier after pass 0; pass 1 -> 2 stays inside a warp (a pair of warps for
//             L = 2048: named barrier).
//   CTAs are persistent and walk work items (a column's chunk of frames) as one continuous frame
//   pipeline, so Mode R (one frame per column) runs at the same rate as long integrations.
// Index algebra restated in numpy and checked on the CPU by tests/test_fft_plan.py.
#pragma once
#include "sti_common.cuh"
#include "r32_math.cuh"
#include <type_traits>

struct R32Args {
    StiArgs s;    // tw = full table W_N^m; chunk / nsplit as in the other kernels
    int nitems;   // ncol * nsub * nsplit
    int ngroups;  // clusters (CTAs for CL = 1) in the grid
    long long* trace;  // R32_TRACE: [frames][16 warps][events] clock64 of CTA 0, else unused
};

// ---- tensor memory as a scratch file ----------------------------------------------------------------------
// .32x32b: thread i of warp w owns TMEM lane 32 (w % 4) + i; the address is warp-uniform (lane base << 16 | column)
PSG_DEV void tmem_ld8(uint32_t taddr, float* r) {
    uint32_t u[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}
PSG_DEV void tmem_ld16(uint32_t taddr, float* r) {
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
          "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}
PSG_DEV void tmem_st8(uint32_t taddr, const float* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])), "r"(__float_as_uint(r[3])),
                 "r"(__float_as_uint(r[4])), "r"(__float_as_uint(r[5])), "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7]))
                 : "memory");
}
PSG_DEV void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- geometry -------------------------------------------------------------------------------------------------
// OPT bit 0: the warps issue their pass-1 loads in groups of four (one warp per scheduler) behind the CTA barrier;
//     bit 1: accumulators as (sum re^2, sum im^2) pairs: one FFMA2 per bin instead of FMUL + FFMA + FADD
//     bit 2: TRACE -- CTA 0 records clock64 at the phase boundaries of frames 8..11 (tools/r32_trace.py)
enum { R32_ORDER = 1, R32_ACC2 = 2, R32_TRACE = 4 };
enum { R32_TRACE_EVENTS = 12, R32_TRACE_FRAME0 = 8, R32_TRACE_FRAMES = 4 };
template <int LOGN, int T_, int IQT, int OPT = 0>
struct R32Cfg {
    using G = R32Geo<LOGN, T_>;
    static constexpr int T = T_, NW = T / 32;
    static constexpr int IQB = IqBytes<IQT>::value;
    static constexpr int NSEG_S = (IQB == 8) ? 24 : 32;  // segments staged in S; the rest by LDG
    static constexpr int NLDG = 32 - NSEG_S;
    static constexpr int SEG = T * IQB + (G::CL == 1 ? 0 : 16);  // staged segment (+ alignment slack; CL = 1: one contiguous copy)
    static constexpr int HDR = 128;
    static constexpr int MBYTES = T * 32 * 8;
    static constexpr int SBYTES = NSEG_S * SEG + (G::CL == 1 ? 128 : 0);  // CL = 1: slack once, rounded so that M stays 128-byte aligned
    static constexpr size_t smem_bytes = HDR + (size_t)SBYTES + MBYTES;
    static_assert(SBYTES % 128 == 0, "M is 128-byte aligned");
    static_assert((size_t)G::NR * G::RS * 4 <= MBYTES, "epilogue staging fits M");
    // TMEM columns of a warp's slot (128 per slot, slot = warp / 4)
    static constexpr int ACCW = (OPT & R32_ACC2) ? 64 : 32;
    static constexpr int C_ACC = 0, C_WIN = ACCW, C_PW0 = ACCW + 32, C_PW1 = ACCW + 48;
};

PSG_DEV cf lds_cf(uint32_t saddr) {
    cf v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
    return v;
}
PSG_DEV void sts_cf(uint32_t saddr, cf v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(saddr), "f"(v.x), "f"(v.y) : "memory"); }
PSG_DEV void lds_cf2(uint32_t saddr, cf& a, cf& b) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(b.x), "=f"(b.y) : "r"(saddr));
}
PSG_DEV void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
// barrier ids as immediates: with an id in a register ptxas reserves all 16 barriers for the CTA, and two CTAs do
// not fit one SM any more (measured: the 8192-point form ran one CTA per SM)
template <int ID, int NTHREADS>
PSG_DEV void named_bar_arrive_c() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(NTHREADS) : "memory"); }
template <int ID, int NTHREADS>
PSG_DEV void named_bar_sync_c() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NTHREADS) : "memory"); }
PSG_DEV void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <int LOGN, int T_, int IQT, int OPT = 0>
__global__ void __launch_bounds__(T_, 512 / T_) sti_r32_kernel(const R32Args ra) {
    using CF = R32Cfg<LOGN, T_, IQT, OPT>;
    using G = typename CF::G;
    constexpr bool ORDER = (OPT & R32_ORDER) != 0, ACC2 = (OPT & R32_ACC2) != 0, TRACE = (OPT & R32_TRACE) != 0;
    constexpr int T = CF::T, NW = CF::NW, N = G::N, L = G::L, CL = G::CL, NR = G::NR, R1 = G::R1, S1 = G::S1, NB1 = G::NB1,
                  IQB = CF::IQB, NSEG_S = CF::NSEG_S, NLDG = CF::NLDG, SEG = CF::SEG, RS = G::RS;
    const StiArgs& a = ra.s;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* const bar_full = reinterpret_cast<uint64_t*>(smem_raw);        // S holds the next frame
    uint64_t* const bar_mfree = reinterpret_cast<uint64_t*>(smem_raw + 8);   // CL == 1: every warp is done with M
    uint64_t* const bar_landed = reinterpret_cast<uint64_t*>(smem_raw + 16);  // CL > 1: the peers' pass-0 outputs (bytes)
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 32);
    unsigned char* const stage = smem_raw + CF::HDR;
    const uint32_t m_base = smem_u32(smem_raw + CF::HDR + (size_t)CF::SBYTES);

    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int c = (CL > 1) ? (int)cluster_ctarank() : 0;
    const int group = blockIdx.x / CL;
    const int np = c * T + t;  // this thread's n' in pass 0

    // ---- work items: group g walks items g, g + ngroups, ...; one continuous sequence of frames ----
    struct Cursor {
        long long base;  // element offset of the frame's first sample
        int item, f, nfr;
        bool valid;
    };
    auto open_item = [&](Cursor& cu, int item) {
        cu.item = item;
        cu.valid = item < ra.nitems;
        cu.f = 0;
        cu.nfr = 0;
        cu.base = 0;
        if (cu.valid) {
            const int split = item % a.nsplit, cs = item / a.nsplit;
            const int col = cs % a.ncol, sub = cs / a.ncol;
            const int kf0 = split * a.chunk;
            cu.nfr = min(a.nfr, kf0 + a.chunk) - kf0;
            cu.base = __ldg(a.col_off + col) + (long long)sub * a.sub_stride + (long long)kf0 * a.hop_elems;
        }
    };
    auto advance = [&](const Cursor& cu) {
        Cursor n = cu;
        if (cu.f + 1 < cu.nfr) {
            n.f = cu.f + 1;
            n.base = cu.base + a.hop_elems;
        } else {
            open_item(n, cu.item + ra.ngroups);
        }
        return n;
    };
    // bulk copies of a frame's staged segments + L2 prefetch of the segments that go through registers.
    // CL = 1: the CTA's segments are the frame itself, contiguous: one bulk copy and one prefetch by thread 0.
    // CL > 1: segments of T samples every L; issuing a bulk copy costs the issuing warp ~70 clk (tools/r32_trace.py:
    // one thread issuing all 32 delayed its warp by 2000 clk a frame), so lane 0 of warp w issues segments w,
    // w + NW, ..; warp 0 also arms the barrier.
    auto issue = [&](long long base) {
        const uintptr_t src0 = reinterpret_cast<uintptr_t>(a.iq) + (uintptr_t)((base + c * T) * IQB);
        const uint32_t slack = (src0 & 15) ? 16 : 0;
        const uintptr_t al = src0 & ~(uintptr_t)15;
        if constexpr (CL == 1) {
            if (t == 0) {
                const uint32_t bytes = NSEG_S * T * IQB + slack;
                mbar_expect_tx(bar_full, bytes);
                bulk_g2s(stage, reinterpret_cast<const void*>(al), bytes, bar_full);
                if constexpr (NLDG > 0)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(al + (uintptr_t)NSEG_S * T * IQB),
                                 "r"((uint32_t)(NLDG * T * IQB) + slack)
                                 : "memory");
            }
        } else if (lane == 0) {
            const uint32_t bytes = T * IQB + slack;
            if (w == 0) mbar_expect_tx(bar_full, bytes * NSEG_S);
#pragma unroll
            for (int h = 0; h < 32 / NW; ++h) {
                const int s = w + NW * h;
                if (s < NSEG_S) bulk_g2s(stage + s * SEG, reinterpret_cast<const void*>(al + (uintptr_t)s * L * IQB), bytes, bar_full);
                else asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(al + (uintptr_t)s * L * IQB), "r"(bytes) : "memory");
            }
        }
    };
    cf pre[NLDG > 0 ? NLDG : 1];  // the frame's last NLDG segments, loaded a pass ahead
    auto load_pre = [&](const Cursor& cu) {
        if constexpr (NLDG > 0) {
#pragma unroll
            for (int i = 0; i < NLDG; ++i)
                pre[i] = cu.valid ? ldg_iq<IQT>(a.iq, cu.base + np + (long long)(NSEG_S + i) * L) : make_float2(0.f, 0.f);
        }
    };

    Cursor cur;
    open_item(cur, group);
    if (!cur.valid) return;  // uniform over the cluster; nothing allocated yet

    // ---- setup: barriers, tensor memory, per-thread tables -> TMEM ----
    if (t == 0) {
        mbar_init(bar_full, 1);
        mbar_init(bar_mfree, NW);
        mbar_init(bar_landed, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(T));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(128 * (w >> 2));
    constexpr uint32_t PEER_BYTES = (uint32_t)(CL - 1) * (uint32_t)T * NR * 8u;  // (CL-1) peers x T columns x NR rows
    if (t == 0) {
        if constexpr (CL > 1) mbar_expect_tx(bar_landed, PEER_BYTES);  // frame 0
    }
    issue(cur.base);
    load_pre(cur);
    {
        // window of this thread's samples n' + a L, a < 32: column 2 j = element j, column 2 j + 1 = element j + 16
        float wv[8];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = 4 * m + (i >> 1), aa = j + 16 * (i & 1);
                wv[i] = __ldg(a.win + np + aa * L);
            }
            tmem_st8(tmem + CF::C_WIN + 8 * m, wv);
        }
        // twiddle powers: pass 0  W_N^{n' 2^q},  pass 1  W_L^{c1 2^q} with c1 = t mod S1 (the same for every butterfly of the thread)
        const int c1 = t & (S1 - 1);
        float p0[16], p1[16];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const float2 v0 = __ldg(a.tw + (((unsigned)np << q) & (N - 1)));
            const float2 v1 = __ldg(a.tw + ((((unsigned)c1 << q) * 32u) & (N - 1)));
            p0[2 * q] = v0.x; p0[2 * q + 1] = v0.y;
            p1[2 * q] = v1.x; p1[2 * q + 1] = v1.y;
        }
#pragma unroll
        for (int i = 10; i < 16; ++i) p0[i] = p1[i] = 0.f;
        tmem_st8(tmem + CF::C_PW0, p0);
        tmem_st8(tmem + CF::C_PW0 + 8, p0 + 8);
        tmem_st8(tmem + CF::C_PW1, p1);
        tmem_st8(tmem + CF::C_PW1 + 8, p1 + 8);
        tmem_wait_st();
    }
    // pass-0 destinations: this thread's column n' of every row; the peers' copies of M and of "landed"
    const uint32_t np_off = r32_p0_col<G>(np);  // row r adds r * L * 8 (a multiple of 1024: the swizzle bits of n' stay)
    uint32_t rbase[CL], rbar[CL];
#pragma unroll
    for (int s = 0; s < CL; ++s) {
        rbase[s] = (CL > 1) ? map_cluster(m_base + np_off, (unsigned)s) : m_base + np_off;
        rbar[s] = (CL > 1) ? map_cluster(smem_u32(bar_landed), (unsigned)s) : 0u;
    }
    if constexpr (CL > 1) {
        __syncthreads();   // "landed" armed before the peers are released
        cluster_arrive();  // matches the wait before the first store of frame 0
    }

    uint32_t q = 0;        // frames done by this CTA
    bool full_ok = false;  // "S full" of the coming frame already observed (probed under the previous frame's last pass)
    // TRACE: x0 / x1 make the clock read depend on the arithmetic before it (the packed-math asm is not volatile)
    auto mark = [&](int ev, float x0, float x1) {
        if constexpr (TRACE) {
            asm volatile("" ::"f"(x0), "f"(x1) : "memory");
            if (blockIdx.x == 0 && lane == 0 && q >= R32_TRACE_FRAME0 && q < R32_TRACE_FRAME0 + R32_TRACE_FRAMES)
                ra.trace[((q - R32_TRACE_FRAME0) * 16 + w) * R32_TRACE_EVENTS + ev] = clock64();
        }
    };
    for (;; ++q) {
        const Cursor nxt = advance(cur);
        mark(0, 0.f, 0.f);
        const int skew = (int)(((reinterpret_cast<uintptr_t>(a.iq) + (uintptr_t)((cur.base + c * T) * IQB)) & 15) / IQB);
        cf x[32];
        bool mfree_ok = true;
        // ---- pass 0: samples -> registers, window folded into the first butterfly layer ----
        if (!full_ok) mbar_wait_bounded(bar_full, q & 1);
        mark(1, 0.f, 0.f);
        // in the order the first butterfly layer consumes them: (j, j + 16)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            x[j] = lds_iq<IQT>(stage + j * SEG, skew + t);
            if (j + 16 < NSEG_S) x[j + 16] = lds_iq<IQT>(stage + (j + 16) * SEG, skew + t);
            else x[j + 16] = pre[j + 16 - NSEG_S];
        }
        {
            cf u[16], v[16];
            float wv[16];
            tmem_ld16(tmem + CF::C_WIN, wv);
            dft32_layer8<0, true>(x, wv, u, v);
            tmem_ld16(tmem + CF::C_WIN + 16, wv);
            dft32_layer8<8, true>(x, wv, u, v);
            mark(2, u[0].x, v[15].y);
            // probe "M free" now: the answer arrives under the sixteen-point butterflies
            if constexpr (CL == 1) mfree_ok = q == 0 || mbar_test(bar_mfree, (q - 1) & 1);
            dft32_finish(x, u, v);
        }
        mark(3, x[1].x, x[31].y);
        // M free: every warp (every CTA of the cluster) is past the last pass of the previous frame
        if constexpr (CL > 1) {
            cluster_wait();
        } else {
            if (!mfree_ok) mbar_wait_bounded(bar_mfree, (q - 1) & 1);
        }
        mark(4, 0.f, 0.f);
        {
            // twiddle W_N^{n' k0} and store, output by output: row k0 % NR of CTA k0 / NR
            float pf[16];
            tmem_ld16(tmem + CF::C_PW0, pf);
            cf pw[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) pw[i] = make_float2(pf[2 * i], pf[2 * i + 1]);
            twiddle_dfs32(x, pw, [&](int k0, cf v) {
                const int s = k0 / NR, r = k0 % NR;
                const uint32_t off = (uint32_t)r * (L * 8);
                if (CL == 1 || s == c) sts_cf(rbase[CL == 1 ? 0 : s] + off, v);
                else st_async_cf(rbase[s] + off, v, rbar[s]);
            });
        }
        mark(5, 0.f, 0.f);
        __syncthreads();  // this CTA's own pass-0 stores
        if constexpr (CL > 1) {
            mbar_wait_bounded(bar_landed, q & 1);  // the peers' (st.async, counted in bytes)
            if (t == 0 && nxt.valid) mbar_expect_tx(bar_landed, PEER_BYTES);  // next frame: sent only after the cluster barrier
        }
        mark(6, 0.f, 0.f);
        // ---- pass 1: NB1 radix-R1 butterflies at stride S1, in place ----
        {
            // ORDER: behind the CTA barrier every warp wants the LSU at once and all of them get their data
            // last; chained named barriers let warps 4 g .. 4 g + 3 (one per scheduler) issue their loads
            // before group g + 1 does, so group 0 computes while the others still load
            if constexpr (ORDER) {
                const int g = w >> 2;
                if (g == 1) named_bar_sync_c<1, 256>();
                if (NW > 8 && g == 2) named_bar_sync_c<2, 256>();
                if (NW > 8 && g == 3) named_bar_sync_c<3, 256>();
            }
            if constexpr (R1 == 32) {
                const uint32_t base1 = m_base + r32_p1_base<G>(t, 0);
#pragma unroll
                for (int j = 0; j < 16; ++j) {  // pair order of the first butterfly layer
                    x[j] = lds_cf(base1 + r32_p1_off<G>(t, j));
                    x[j + 16] = lds_cf(base1 + r32_p1_off<G>(t, j + 16));
                }
            } else {
#pragma unroll
                for (int i = 0; i < NB1; ++i) {
                    const uint32_t base1 = m_base + r32_p1_base<G>(t, i);
#pragma unroll
                    for (int j = 0; j < 4; ++j)  // the order of the first layer of 4-point butterflies
#pragma unroll
                        for (int h = 0; h < 4; ++h) x[16 * i + j + 4 * h] = lds_cf(base1 + r32_p1_off<G>(t, j + 4 * h));
                }
            }
            if constexpr (ORDER) {
                const int g = w >> 2;
                if (g == 0) named_bar_arrive_c<1, 256>();
                if (NW > 8 && g == 1) named_bar_arrive_c<2, 256>();
                if (NW > 8 && g == 2) named_bar_arrive_c<3, 256>();
            }
            // S was consumed by every warp before the barrier.  It is refilled from here, not right after its last
            // read: the bulk copies then land under the arithmetic of passes 1 and 2 instead of competing with the
            // pass-0 stores and pass-1 loads, the one stretch of the frame that is bound by the shared-memory pipe
            if (nxt.valid) issue(nxt.base);
            mark(7, x[0].x, x[31].y);
            float pf[16];
            tmem_ld16(tmem + CF::C_PW1, pf);
            cf pw[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) pw[i] = make_float2(pf[2 * i], pf[2 * i + 1]);
            if constexpr (R1 == 32) {
                const uint32_t base1 = m_base + r32_p1_base<G>(t, 0);
                dft32(x);
                twiddle_dfs32(x, pw, [&](int k1, cf v) { sts_cf(base1 + r32_p1_off<G>(t, k1), v); });
            } else {
#pragma unroll
                for (int i = 0; i < NB1; ++i) {
                    const uint32_t base1 = m_base + r32_p1_base<G>(t, i);
                    dft16(&x[16 * i]);
                    twiddle_dfs16(&x[16 * i], pw, [&](int k1, cf v) { sts_cf(base1 + r32_p1_off<G>(t, k1), v); });
                }
            }
            mark(8, 0.f, 0.f);
        }
        if constexpr (S1 == 64) named_bar_sync(4 + (t >> 6), 64);  // a row is two warps (ids 4..11; one CTA per SM here)
        else __syncwarp();                                        // a warp owns whole rows
        mark(9, 0.f, 0.f);
        // next frame's register segments: in flight under pass 2
        load_pre(nxt);
        full_ok = nxt.valid && mbar_test(bar_full, (q + 1) & 1);
        // ---- pass 2: the S1-point DFTs on consecutive positions, |X|^2 into the accumulators (TMEM) ----
        const bool first = cur.f == 0;
        tmem_wait_st();  // the accumulator columns written by the previous frame
        // bins bin0 .. bin0 + NB - 1 of this thread += |y|^2
        auto accumulate = [&](int bin0, const cf* y, auto nb_tag) {
            constexpr int NB = decltype(nb_tag)::value;
            if constexpr (ACC2) {
#pragma unroll
                for (int g = 0; g < NB / 4; ++g) {
                    float acc[8];
                    const uint32_t col = tmem + CF::C_ACC + 2 * bin0 + 8 * g;
                    if (first) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const cf sq = mul2(y[4 * g + i], y[4 * g + i]);
                            acc[2 * i] = sq.x;
                            acc[2 * i + 1] = sq.y;
                        }
                    } else {
                        tmem_ld8(col, acc);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const cf sq = fma2(y[4 * g + i], y[4 * g + i], make_float2(acc[2 * i], acc[2 * i + 1]));
                            acc[2 * i] = sq.x;
                            acc[2 * i + 1] = sq.y;
                        }
                    }
                    tmem_st8(col, acc);
                }
            } else {
#pragma unroll
                for (int g = 0; g < NB / 8; ++g) {
                    float acc[8];
                    const uint32_t col = tmem + CF::C_ACC + bin0 + 8 * g;
                    if (first) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] = fmaf(y[8 * g + i].x, y[8 * g + i].x, y[8 * g + i].y * y[8 * g + i].y);
                    } else {
                        tmem_ld8(col, acc);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            acc[i] += fmaf(y[8 * g + i].x, y[8 * g + i].x, y[8 * g + i].y * y[8 * g + i].y);
                    }
                    tmem_st8(col, acc);
                }
            }
        };
        if constexpr (S1 == 16) {
            // two blocks of 16 consecutive elements (one line each) out of the rows this warp wrote in pass 1
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                cf y[16];
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) lds_cf2(m_base + r32_p2_addr<G>(t, i, ch), y[2 * ch], y[2 * ch + 1]);
                dft16(y);
                if (i == 0) mark(10, y[0].x, y[15].y);
                accumulate(16 * i, y, std::integral_constant<int, 16>{});
            }
        } else if constexpr (S1 == 32) {
            // row w, block k1 = lane: 32 consecutive elements = lines 2 lane, 2 lane + 1, chunk (j & 7) ^ (lane & 7)
#pragma unroll
            for (int ch = 0; ch < 16; ++ch) lds_cf2(m_base + r32_p2_addr<G>(t, 0, ch), x[2 * ch], x[2 * ch + 1]);
            dft32(x);
            mark(10, x[0].x, x[31].y);
            accumulate(0, x, std::integral_constant<int, 32>{});
        } else {
            // row r2 = t / 64, k1 = (t % 64) / 2, e = t & 1: elements 64 k1 + 2 d + e, d < 32; then the radix-2
            // butterfly over e between lanes l and l ^ 1: lane e = 0 finishes outputs q2 = i and i + 32, i < 16,
            // lane e = 1 outputs 16 + i and 48 + i (W_64^{16 + i} = -j W_64^i)
            const int e = t & 1;
#pragma unroll
            for (int d = 0; d < 32; ++d) x[d] = lds_cf(m_base + r32_p2_addr<G>(t, 0, d));
            dft32(x);
            mark(10, x[0].x, x[31].y);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                cf s[8];
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const int i = 4 * m + ii;
                    const cf send = e ? x[i] : x[16 + i];
                    const cf keep = e ? x[16 + i] : x[i];
                    cf recv;
                    recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                    recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                    r32_pair_finish(e, i, keep, recv, s[2 * ii], s[2 * ii + 1]);
                }
                accumulate(8 * m, s, std::integral_constant<int, 8>{});
            }
        }
        tmem_wait_st();
        mark(11, 0.f, 0.f);
        // ---- the item's last frame: accumulators -> fftshifted column, coalesced 128-bit stores ----
        if (cur.f + 1 == cur.nfr) {
            __syncthreads();  // every warp is done reading M (peers write to it only after the cluster barrier)
            float* const sout = reinterpret_cast<float*>(smem_raw + CF::HDR + (size_t)CF::SBYTES);
            // this CTA's bins: freq = k0 + 32 m, k0 = c NR + r; fftshift moves m by N/64; staged as sout[r RS + m']
            constexpr int MM = N / 32, MSH = N / 64;
#pragma unroll
            for (int m8 = 0; m8 < CF::ACCW / 8; ++m8) {
                float acc[8];
                tmem_ld8(tmem + CF::C_ACC + 8 * m8, acc);
#pragma unroll
                for (int j = 0; j < (ACC2 ? 4 : 8); ++j) {
                    const int ai = (ACC2 ? 4 : 8) * m8 + j;  // accumulator index -> (row r, m)
                    int r, m;
                    r32_acc_bin<G>(t, ai, r, m);
                    sout[r * RS + ((m + MSH) & (MM - 1))] = ACC2 ? acc[2 * j] + acc[2 * j + 1] : acc[j];
                }
            }
            __syncthreads();
            const int cs = cur.item / a.nsplit, split = cur.item % a.nsplit;
            constexpr int QPM = NR / 4;  // float4 per m'
            for (int e4 = t; e4 < MM * QPM; e4 += T) {
                const int mp = e4 / QPM, r0 = 4 * (e4 % QPM);
                float4 v4 = make_float4(sout[(r0 + 0) * RS + mp], sout[(r0 + 1) * RS + mp], sout[(r0 + 2) * RS + mp],
                                        sout[(r0 + 3) * RS + mp]);
                const size_t o4 = (size_t)(32 * mp + c * NR + r0) / 4;  // float4 index inside the column
                if (a.nsplit > 1) {
                    reinterpret_cast<float4*>(a.partial + ((size_t)cs * a.nsplit + split) * N)[o4] = v4;
                } else {
                    v4.x *= a.scale; v4.y *= a.scale; v4.z *= a.scale; v4.w *= a.scale;
                    const size_t o = (size_t)cs * (N / 4) + o4;
                    if (a.out_lin) reinterpret_cast<float4*>(a.out_lin)[o] = v4;
                    if (a.out_db)
                        reinterpret_cast<float4*>(a.out_db)[o] = make_float4(power_to_db(v4.x, a.eps), power_to_db(v4.y, a.eps),
                                                                             power_to_db(v4.z, a.eps), power_to_db(v4.w, a.eps));
                }
            }
        }
        // ---- done with M ----
        if constexpr (CL > 1) {
            cluster_arrive_relaxed();
        } else {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_mfree);
        }
        if (!nxt.valid) break;
        cur = nxt;
    }
    if constexpr (CL > 1) cluster_wait();  // consume the last arrive; no peer writes to this CTA any more
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "n"(T));
}
