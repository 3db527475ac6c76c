// nfft = R0 * 4096 (R0 = 2, 4: 8192 and 16384 points) with the WHOLE frame in one SM's shared memory.
//
// A 16384-point complex64 frame is 128 KB: its padded exchange buffer (144 KB) fits the 227 KB of one
// CTA, its staging ring does not if it has to hold whole frames.  So the first pass is fed in slabs:
//   plan      N = R0 x 16 x 16 x 16, strides S = 4096, 256, 16, 1 (the index algebra of sti_kernels.cuh)
//   pass 0    radix R0 over n0 (elements n' + n0*4096).  The slab n' in [m*512, (m+1)*512) arrives as R0
//             bulk copies (UBLKCP) of 4 KB into one stage of a ring; thread t owns n' = m*512 + t, reads
//             its R0 samples, does the windowed R0-point DFT, multiplies by W_N^{n' k0} and stores output
//             k0 to position k0*4096 + n' of the exchange buffer.  W_N^{n'} = W_N^{t} * W_N^{512 m}: the
//             first factor is a per-thread register, the second a kernel-argument constant, so the
//             pre-pass loads no twiddles; the window (w/sum(w), 4 B per sample) is read with coalesced
//             LDG before the stage is waited for.
//   passes 1-3  in place on the whole frame, radix 16, N/8192 butterflies per thread; mid-pass twiddles
//             are rebuilt from W^1,2,4,8 in registers (the tables of the 4096-point plan); pass 2 -> 3
//             stays inside aligned groups of 16 threads (__syncwarp); the last pass accumulates |X|^2.
// 512 threads, one CTA per SM, three CTA barriers per frame.  The ring is refilled by whichever warp
// reads a stage last (a shared-memory counter per stage), so nobody ever waits for "stage free": the
// slabs of frame f+1 stream in under passes 1-3 of frame f.  (A dedicated producer warp would make the
// CTA 17 warps, which the register file allocates as 20: 96 registers per thread and spills.)
// Everything is local to the SM -- no
// cluster, no remote shared memory, no global scratch -- which is what the cluster kernels
// (sti_cluster.cuh) pay for: same four shared-memory round trips per sample, but 16 warps that never
// wait on another SM.  HBM sees every sample once.
#pragma once
#include "sti_cluster.cuh"

struct WholeArgs {
    StiArgs s;        // tw = full table W_N^m, twp = pass tables of the 4096-point 16x16x16 plan (power layout)
    float2 cm[4][16];  // W_N^{256 * j * 2^q}, j < 16, q < 4
};

template <int R0, int IQT, int NST>
struct WholeCfg {
    static constexpr int N2 = 4096, T = 512, W = 512;
    static constexpr int N = R0 * N2;
    static constexpr int NBT = N / (16 * T);  // radix-16 butterflies per thread and pass
    static constexpr int NSLAB = N2 / W;      // 8 slabs per frame
    static constexpr int IQB = IqBytes<IQT>::value;
    static constexpr int SEG = W * IQB + 16;  // staged segment + alignment slack
    static constexpr int STAGE = R0 * SEG;
    static constexpr int NPAD = psg_pad(N) + 2;
    static constexpr int HDR = 128;  // NST mbarriers + NST reader counters
    static constexpr size_t smem_bytes = HDR + (size_t)NST * STAGE + (size_t)NPAD * 8;
    static_assert(NSLAB % NST == 0 && NST <= 8, "the stage of a slab is a compile-time constant");
};


template <int R0, int IQT, int NST>
__global__ void __launch_bounds__(512, 1) sti_whole_kernel(const WholeArgs wa) {
    using CF = WholeCfg<R0, IQT, NST>;
    constexpr int N2 = CF::N2, N = CF::N, T = CF::T, W = CF::W, NBT = CF::NBT, NSLAB = CF::NSLAB, IQB = CF::IQB, SEG = CF::SEG;
    constexpr int NW = T / 32;
    constexpr int NPWA = psg_npow(R0);
    using PL4 = Plan<N2, 16, 16, 16, 1, 2>;  // table offsets of the 4096-point plan
    using PLN = Plan<N, R0, 16, 16, 16, 2>;  // frequency map of the whole transform
    static_assert(PL4::ROW1 && PL4::S0 == 256 && PL4::S1 == 16, "4096 = 16*16*16 plan");
    const StiArgs& a = wa.s;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);          // [NST] stage full
    unsigned* cnt = reinterpret_cast<unsigned*>(smem_raw + 64);      // [NST] warps that have read the stage (running total)
    unsigned char* stage = smem_raw + CF::HDR;
    float2* xch = reinterpret_cast<float2*>(smem_raw + CF::HDR + (size_t)NST * CF::STAGE);

    const int t = threadIdx.x;
    const int item = blockIdx.x;
    const int split = item % a.nsplit;
    const int cs = item / a.nsplit;
    const int col = cs % a.ncol, sub = cs / a.ncol;
    const int kf0 = split * a.chunk;
    const int nfr = min(a.nfr, kf0 + a.chunk) - kf0;
    const long long fbase = a.col_off[col] + (long long)sub * a.sub_stride + (long long)kf0 * a.hop_elems;
    const int nsteps = nfr * NSLAB;  // step q = (frame q / NSLAB, slab q % NSLAB) -> stage q % NST

    auto issue = [&](int q) {  // one thread
        const int f = q / NSLAB, m = q % NSLAB, s = q % NST;
        const uintptr_t src0 =
            reinterpret_cast<uintptr_t>(a.iq) + (uintptr_t)((fbase + (long long)f * a.hop_elems + m * W) * IQB);
        const uint32_t bytes = W * IQB + ((src0 & 15) ? 16 : 0);
        uint64_t* bar = bars + s;
        mbar_expect_tx(bar, bytes * R0);
        unsigned char* dst = stage + (size_t)s * CF::STAGE;
#pragma unroll
        for (int n0 = 0; n0 < R0; ++n0)
            bulk_g2s(dst + n0 * SEG, reinterpret_cast<const void*>((src0 & ~(uintptr_t)15) + (uintptr_t)n0 * N2 * IQB), bytes, bar);
    };
    if (t == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bars + s, 1);
            cnt[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int q = 0; q < NST && q < nsteps; ++q) issue(q);
    }

    // loop-invariant registers: W_N^{t 2^q}, the twiddle powers of passes 1 and 2
    cf wt[NPWA];
#pragma unroll
    for (int q = 0; q < NPWA; ++q) wt[q] = __ldg(a.tw + (t << q));
    // (the twiddle powers of passes 1 and 2 are re-read from L1 before the barrier that precedes the pass:
    // as loop-invariant registers they push the accumulators into local memory)
    const float2* const tw1p = a.twp + PL4::TW0 + (t & 255);     // W_4096^{(t & 255) 2^q} at [((1 << q) - 1) * 256]
    const float2* const tw2p = a.twp + PL4::TW1 + (t & 15) * 6;  // W_256^{(t & 15) 2^q} at [q]
    float acc[16 * NBT];
#pragma unroll
    for (int i = 0; i < 16 * NBT; ++i) acc[i] = 0.f;
    __syncthreads();  // barriers and counters initialised

    float2* const p0 = xch + psg_pad(t);
    // window of this thread's samples: the table (4 B per sample, up to 64 KB) does not fit what is
    // left of L1 beside 213 KB of shared memory, so every load is an L2 round trip (~700 clk) and has to
    // be in flight long before it is used: slabs 0..3 of the next frame are requested before the last
    // pass of the current one, slabs 4,5 / 6,7 two slab pairs ahead of their use inside pass 0.
    constexpr int WQ = 4;
    float wq[2 * WQ][R0];
    auto load_window = [&](int m, float* w) {
#pragma unroll
        for (int n0 = 0; n0 < R0; ++n0) w[n0] = __ldg(a.win + n0 * N2 + m * W + t);
    };
#pragma unroll
    for (int m = 0; m < WQ; ++m) load_window(m, wq[m]);
    for (int f = 0; f < nfr; ++f) {
        const int skew =
            (int)(((reinterpret_cast<uintptr_t>(a.iq) + (uintptr_t)((fbase + (long long)f * a.hop_elems) * IQB)) & 15) / IQB);
        load_window(4, wq[4]);
        load_window(5, wq[5]);
        cf wb1[4];  // pass-1 twiddle powers (requested half a pass ahead: the window loads keep evicting them from L1)
        // ---- pass 0, two slabs at a time ----
#pragma unroll
        for (int mp = 0; mp < NSLAB / 2; ++mp) {
            cf x[2][R0];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = 2 * mp + h, s = m % NST, q = f * NSLAB + m;
                mbar_wait(bars + s, (q / NST) & 1);
                const unsigned char* sb = stage + (size_t)s * CF::STAGE;
#pragma unroll
                for (int n0 = 0; n0 < R0; ++n0) x[h][n0] = lds_iq<IQT>(sb + n0 * SEG, skew + t);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = 2 * mp + h;
                dftRw<R0>(x[h], wq[m]);
            }
            if (mp == 0) {
                load_window(6, wq[6]);
                load_window(7, wq[7]);
            }
            if (mp == 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) wb1[q] = __ldg(tw1p + ((1 << q) - 1) * PL4::S0);
            }
            // the butterflies have consumed every load of this warp: the warp that reads a stage last refills it
            __syncwarp();
            if ((t & 31) == 0) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int m = 2 * mp + h, s = m % NST, q = f * NSLAB + m;
                    const unsigned old = count_reader(cnt + s);
                    if ((old & (NW - 1)) == NW - 1 && q + NST < nsteps) issue(q + NST);
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = 2 * mp + h;
                cf pw[NPWA];
#pragma unroll
                for (int qq = 0; qq < NPWA; ++qq) pw[qq] = (m == 0) ? wt[qq] : cmul(wt[qq], wa.cm[qq][2 * m]);
                twiddle_dfs<R0>(x[h], pw);
            }
            if (mp == 0) __syncthreads();  // the last pass of the previous frame is done with the exchange buffer
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int k0 = 0; k0 < R0; ++k0) p0[pad_off(k0 * N2 + (2 * mp + h) * W)] = x[h][k0];
        }
        __syncthreads();
        // ---- pass 1: radix 16, stride 256, inside each 4096-point row ----
        // (butterflies of a thread one after the other: interleaved they spill at 128 registers)
#pragma unroll 1
        for (int i = 0; i < NBT; ++i) {
            const int row = (t + i * T) >> 8;
            float2* p = xch + psg_pad(t & 255) + pad_off(row * N2);
            cf v[16];
#pragma unroll
            for (int n = 0; n < 16; ++n) v[n] = p[pad_off(n * 256)];
            dftR<16>(v);
            twiddle_dfs<16>(v, wb1);
#pragma unroll
            for (int k = 0; k < 16; ++k) p[pad_off(k * 256)] = v[k];
        }
        cf wb2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) wb2[q] = __ldg(tw2p + q);
        __syncthreads();
        // ---- pass 2: radix 16, stride 16, inside each 256-point block ----
#pragma unroll 1
        for (int i = 0; i < NBT; ++i) {
            float2* p = xch + psg_pad(((t + i * T) >> 4) * 256) + (t & 15);
            cf v[16];
#pragma unroll
            for (int n = 0; n < 16; ++n) v[n] = p[pad_off(n * 16)];
            dftR<16>(v);
            twiddle_dfs<16>(v, wb2);
#pragma unroll
            for (int k = 0; k < 16; ++k) p[pad_off(k * 16)] = v[k];
        }
        __syncwarp();  // the 256-point blocks of passes 2 and 3 stay inside aligned groups of 16 threads
#pragma unroll
        for (int m = 0; m < WQ; ++m) load_window(m, wq[m]);  // next frame's first slabs, in flight under pass 3
        // ---- pass 3: radix 16 on consecutive positions, |X|^2 into the accumulators ----
        smem_pass<16 * NBT, T, 16, 1, true>(xch, nullptr, t, acc);
    }

    // ---- epilogue: digit-reversed register sums -> fftshifted, coalesced 128-bit stores ----
    __syncthreads();
    float* sout = reinterpret_cast<float*>(xch);  // N floats
#pragma unroll
    for (int i = 0; i < NBT; ++i) {
        const int klow = PLN::low_freq(t + i * T);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int freq = klow + (N / 16) * jj;
            const int idx = (freq + N / 2) & (N - 1);
            sout[idx ^ (((idx >> 5) & 7) << 2)] = acc[i * 16 + jj];
        }
    }
    __syncthreads();
    const float4* sout4 = reinterpret_cast<const float4*>(sout);
    constexpr int NQ = N / 4;
    for (int q = t; q < NQ; q += T) {
        float4 v = sout4[q ^ ((q >> 3) & 7)];
        if (a.nsplit > 1) {
            reinterpret_cast<float4*>(a.partial + ((size_t)cs * a.nsplit + split) * N)[q] = v;
        } else {
            v.x *= a.scale; v.y *= a.scale; v.z *= a.scale; v.w *= a.scale;
            const size_t o = (size_t)cs * NQ + q;
            if (a.out_lin) reinterpret_cast<float4*>(a.out_lin)[o] = v;
            if (a.out_db)
                reinterpret_cast<float4*>(a.out_db)[o] =
                    make_float4(power_to_db(v.x, a.eps), power_to_db(v.y, a.eps), power_to_db(v.z, a.eps), power_to_db(v.w, a.eps));
        }
    }
}

// ---- the same kernel on a thread-block cluster: CL CTAs of ROWS 4096-point rows each ------------------------
// N = R0 x 4096 with R0 = ROWS CL = 4 NB.  Split the first-pass index n0 = NB a + b (a < 4, b < NB) and the row
// index k0 = r + 4 s (r < 4, s < NB):
//   X[r + 4 s] = sum_b W_NB^{b s} * W_R0^{b r} * ( sum_a w x[(NB a + b) 4096 + n'] W_4^{a r} )
// CTA c of the cluster owns rows c ROWS .. c ROWS + ROWS - 1 in its exchange buffer and runs passes 1-3 on
// them exactly like the single-CTA kernel.  Its share of the first pass is n' in [c 4096/CL, (c+1) 4096/CL), in
// eight steps (slab m of T columns, then b): a step stages the four samples a = 0..3 of one b (the stages of
// the single-CTA kernel), does the windowed 4-point DFT over a, applies the constant W_R0^{b r} and adds into
// the thread's R0 outputs with the trivial factors W_NB^{b s}; after the last b the outputs get W_N^{n' k0} and
// go to the row owners' buffers, already in the padded layout: ROWS of them with local stores, the rest with
// st.async into the peers' shared memory.
//   ROWS = 4, 512 threads, one CTA per SM:   32768 on 2 CTAs (default), 65536 on 4
//   ROWS = 2, 256 threads, two CTAs per SM:  16384 on 2 CTAs, 32768 on 4, 65536 on 8 -- the two CTAs of an SM
//            belong to different clusters, so the stalls of one first pass hide behind the other's passes
// Synchronisation per frame: one cluster barrier "every CTA is done with the previous frame's last pass"
// (relaxed arrive right after that pass -- nothing is published, so no fence -- wait before the first
// store of the next frame), and for "all first-pass outputs have landed" a CTA barrier for the local
// stores plus an mbarrier that counts the bytes the peers send with st.async (a release / acquire cluster
// barrier here costs a MEMBAR + ERRBAR per thread behind the remote stores: 10 % of all stall samples).
template <int CL, int ROWS, int IQT>
struct WholeClCfg {
    static constexpr int N2 = 4096, T = 128 * ROWS, W = T, NST = 4, NSTEP = 8;
    static constexpr int R0 = ROWS * CL, N = R0 * N2, NB = R0 / 4;
    static constexpr int NPC = N2 / CL;  // n' per CTA
    static constexpr int IQB = IqBytes<IQT>::value;
    static constexpr int SEG = W * IQB + 16;
    static constexpr int STAGE = 4 * SEG;
    static constexpr int NPAD = psg_pad(ROWS * N2) + 2;
    static constexpr int HDR = 128;
    static constexpr size_t smem_bytes = HDR + (size_t)NST * STAGE + (size_t)NPAD * 8;
    static_assert((NPC / W) * NB == NSTEP, "eight steps per frame");
};

// W_16^K = exp(-2 pi j K / 16)
template <int K>
PSG_DEV cf w16_const() {
    constexpr float c[16] = {1.f, PSG_C1_16, PSG_SQRT1_2, PSG_S1_16, 0.f, -PSG_S1_16, -PSG_SQRT1_2, -PSG_C1_16,
                             -1.f, -PSG_C1_16, -PSG_SQRT1_2, -PSG_S1_16, 0.f, PSG_S1_16, PSG_SQRT1_2, PSG_C1_16};
    return make_float2(c[K & 15], -c[(K + 12) & 15]);  // sin(x) = cos(x - pi/2): index K - 4 = K + 12 (mod 16)
}
// acc + v * (-j)^K
template <int K>
PSG_DEV cf add_rot(cf acc, cf v) {
    if constexpr ((K & 3) == 0) return cadd(acc, v);
    else if constexpr ((K & 3) == 1) return cadd(acc, mul_nj(v));
    else if constexpr ((K & 3) == 2) return csub(acc, v);
    else return csub(acc, mul_nj(v));
}

// step b of a slab for one thread: y[r] (the windowed 4-point DFT over a) -> out[r + 4 s] (+)= W_NB^{b s} W_R0^{b r} y[r]
template <int NB, int B>
PSG_DEV void wholec_accumulate(cf* out, cf* y) {
    constexpr int R0 = 4 * NB;
    if constexpr (B > 0 && B < NB) {
        y[1] = cmul(y[1], w16_const<(16 / R0) * B * 1>());
        y[2] = cmul(y[2], w16_const<(16 / R0) * B * 2>());
        y[3] = cmul(y[3], w16_const<(16 / R0) * B * 3>());
    }
    if constexpr (B < NB) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if constexpr (B == 0) {
                out[r] = y[r];
                if constexpr (NB >= 2) out[r + 4] = y[r];
                if constexpr (NB >= 4) { out[r + 8] = y[r]; out[r + 12] = y[r]; }
            } else {
                // W_NB^{b s} = (-j)^{(4 / NB) b s}
                out[r] = cadd(out[r], y[r]);
                if constexpr (NB >= 2) out[r + 4] = add_rot<(4 / NB) * B * 1>(out[r + 4], y[r]);
                if constexpr (NB >= 4) {
                    out[r + 8] = add_rot<(4 / NB) * B * 2>(out[r + 8], y[r]);
                    out[r + 12] = add_rot<(4 / NB) * B * 3>(out[r + 12], y[r]);
                }
            }
        }
    }
}
template <int NB>
PSG_DEV void wholec_accumulate_b(int b, cf* out, cf* y) {  // b is a compile-time constant after unrolling
    switch (b) {
        case 0: wholec_accumulate<NB, 0>(out, y); break;
        case 1: wholec_accumulate<NB, 1>(out, y); break;
        case 2: wholec_accumulate<NB, 2>(out, y); break;
        default: wholec_accumulate<NB, 3>(out, y); break;
    }
}

template <int CL, int ROWS, int IQT>
__global__ void __launch_bounds__(128 * ROWS, 4 / ROWS) sti_wholec_kernel(const WholeArgs wa) {
    using CF = WholeClCfg<CL, ROWS, IQT>;
    constexpr int N2 = CF::N2, N = CF::N, T = CF::T, W = CF::W, R0 = CF::R0, NB = CF::NB, NPC = CF::NPC, NSTEP = CF::NSTEP,
                  NST = CF::NST, IQB = CF::IQB, SEG = CF::SEG;
    constexpr int NW = T / 32, NBT = 2;
    constexpr int NPWA = psg_npow(R0);
    using PL4 = Plan<N2, 16, 16, 16, 1, 2>;
    using PLN = Plan<N, R0, 16, 16, 16, 2>;
    static_assert((ROWS == 4 || ROWS == 2) && CL >= 2 && R0 >= 4 && R0 <= 16, "rows per CTA and cluster size");
    const StiArgs& a = wa.s;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    unsigned* cnt = reinterpret_cast<unsigned*>(smem_raw + 64);
    uint64_t* landed = reinterpret_cast<uint64_t*>(smem_raw + 96);  // bytes of first-pass outputs received from the peers
    unsigned char* stage = smem_raw + CF::HDR;
    float2* xch = reinterpret_cast<float2*>(smem_raw + CF::HDR + (size_t)NST * CF::STAGE);
    constexpr uint32_t PEER_BYTES = (uint32_t)(CL - 1) * NPC * ROWS * 8;  // (CL-1) peers x NPC columns x ROWS rows

    const int t = threadIdx.x;
    const int c = (int)cluster_ctarank();
    const int item = blockIdx.x / CL;
    const int split = item % a.nsplit;
    const int cs = item / a.nsplit;
    const int col = cs % a.ncol, sub = cs / a.ncol;
    const int kf0 = split * a.chunk;
    const int nfr = min(a.nfr, kf0 + a.chunk) - kf0;
    const long long fbase = a.col_off[col] + (long long)sub * a.sub_stride + (long long)kf0 * a.hop_elems + c * NPC;
    const int nsteps = nfr * NSTEP;  // step q = (frame q / 8, slab (q % 8) / NB, b = q % NB) -> stage q % NST

    auto issue = [&](int q) {  // one thread
        const int f = q / NSTEP, e = q % NSTEP, m = e / NB, b = e % NB, s = q % NST;
        const uintptr_t src0 =
            reinterpret_cast<uintptr_t>(a.iq) + (uintptr_t)((fbase + (long long)f * a.hop_elems + b * N2 + m * W) * IQB);
        const uint32_t bytes = W * IQB + ((src0 & 15) ? 16 : 0);
        uint64_t* bar = bars + s;
        mbar_expect_tx(bar, bytes * 4);
        unsigned char* dst = stage + (size_t)s * CF::STAGE;
#pragma unroll
        for (int aa = 0; aa < 4; ++aa)
            bulk_g2s(dst + aa * SEG, reinterpret_cast<const void*>((src0 & ~(uintptr_t)15) + (uintptr_t)aa * NB * N2 * IQB), bytes,
                     bar);
    };
    if (t == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bars + s, 1);
            cnt[s] = 0;
        }
        mbar_init(landed, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(landed, PEER_BYTES);  // frame 0
        for (int q = 0; q < NST && q < nsteps; ++q) issue(q);
    }

    cf wt[NPWA];  // W_N^{(c NPC + t) 2^q}
#pragma unroll
    for (int q = 0; q < NPWA; ++q) wt[q] = __ldg(a.tw + ((c * NPC + t) << q));
    const float2* const tw1p = a.twp + PL4::TW0 + (t & 255);
    const float2* const tw2p = a.twp + PL4::TW1 + (t & 15) * 6;
    float acc[16 * NBT];
#pragma unroll
    for (int i = 0; i < 16 * NBT; ++i) acc[i] = 0.f;

    // this thread's column of the exchange buffers: local pointer and the peers' shared::cluster addresses
    float2* const p0 = xch + psg_pad(t) + pad_off(c * NPC);
    uint32_t rbase[CL], rbar[CL];
#pragma unroll
    for (int s = 0; s < CL; ++s) {
        rbase[s] = map_cluster(smem_u32(p0), (unsigned)s);
        rbar[s] = map_cluster(smem_u32(landed), (unsigned)s);
    }
    __syncthreads();   // barriers and counters initialised, "landed" armed
    cluster_arrive();  // matches the wait before the first store of frame 0 (release: the peers see the armed barrier)

    // window of step e: samples (NB a + b) 4096 + c NPC + T m + t, a = 0..3
    const float* const winp = a.win + c * NPC + t;
    float wq[NSTEP][4];
    auto load_window = [&](int e, float* w) {
        const int m = e / NB, b = e % NB;
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) w[aa] = __ldg(winp + (NB * aa + b) * N2 + m * W);
    };
    // steps requested before the last pass of the previous frame / at the top of the first pass; the rest two
    // pairs ahead of their use.  Sixteen outputs per thread: two steps less in flight (registers)
    constexpr int WQA = (NB == 4) ? 2 : 4, WQB = WQA + 2;
#pragma unroll
    for (int e = 0; e < WQA; ++e) load_window(e, wq[e]);
    for (int f = 0; f < nfr; ++f) {
        const int skew =
            (int)(((reinterpret_cast<uintptr_t>(a.iq) + (uintptr_t)((fbase + (long long)f * a.hop_elems) * IQB)) & 15) / IQB);
#pragma unroll
        for (int e = WQA; e < WQB; ++e) load_window(e, wq[e]);
        cf wb1[4];
        cf out[R0];
        // ---- pass 0, two steps at a time ----
#pragma unroll
        for (int pp = 0; pp < NSTEP / 2; ++pp) {
            cf x[2][4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = 2 * pp + h, s = e % NST, q = f * NSTEP + e;
                mbar_wait(bars + s, (q / NST) & 1);
                const unsigned char* sb = stage + (size_t)s * CF::STAGE;
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) x[h][aa] = lds_iq<IQT>(sb + aa * SEG, skew + t);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) dftRw<4>(x[h], wq[2 * pp + h]);
            if (WQB + 2 * pp < NSTEP) {
                load_window(WQB + 2 * pp, wq[WQB + 2 * pp]);
                load_window(WQB + 2 * pp + 1, wq[WQB + 2 * pp + 1]);
            }
            if (pp == 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) wb1[q] = __ldg(tw1p + ((1 << q) - 1) * PL4::S0);
            }
            __syncwarp();
            if ((t & 31) == 0) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int e = 2 * pp + h, s = e % NST, q = f * NSTEP + e;
                    const unsigned old = count_reader(cnt + s);
                    if ((old & (NW - 1)) == NW - 1 && q + NST < nsteps) issue(q + NST);
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = 2 * pp + h, m = e / NB, b = e % NB;
                wholec_accumulate_b<NB>(b, out, x[h]);
                if (b == NB - 1) {  // slab complete: twiddle, send to the row owners
                    cf pw[NPWA];
#pragma unroll
                    for (int qq = 0; qq < NPWA; ++qq) pw[qq] = (m == 0) ? wt[qq] : cmul(wt[qq], wa.cm[qq][m * (W / 256)]);
                    twiddle_dfs<R0>(out, pw);
                    if (m == 0) cluster_wait();  // every CTA is done with the last pass of the previous frame
#pragma unroll
                    for (int k0 = 0; k0 < R0; ++k0) {
                        const int s = k0 / ROWS, r = k0 % ROWS;
                        const int off = pad_off(r * N2 + m * W);
                        if (s == c) p0[off] = out[k0];
                        else st_async_cf(rbase[s] + (uint32_t)off * 8u, out[k0], rbar[s]);
                    }
                }
            }
        }
        __syncthreads();                    // this CTA's own first-pass stores
        mbar_wait_bounded(landed, f & 1);  // the peers' (st.async, counted in bytes)
        if (t == 0 && f + 1 < nfr) mbar_expect_tx(landed, PEER_BYTES);  // next frame: sent only after the cluster barrier below
        // ---- pass 1 ----
#pragma unroll 1
        for (int i = 0; i < NBT; ++i) {
            const int row = (t + i * T) >> 8;
            float2* p = xch + psg_pad(t & 255) + pad_off(row * N2);
            cf v[16];
#pragma unroll
            for (int n = 0; n < 16; ++n) v[n] = p[pad_off(n * 256)];
            dftR<16>(v);
            twiddle_dfs<16>(v, wb1);
#pragma unroll
            for (int k = 0; k < 16; ++k) p[pad_off(k * 256)] = v[k];
        }
        cf wb2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) wb2[q] = __ldg(tw2p + q);
        __syncthreads();
        // ---- pass 2 ----
#pragma unroll 1
        for (int i = 0; i < NBT; ++i) {
            float2* p = xch + psg_pad(((t + i * T) >> 4) * 256) + (t & 15);
            cf v[16];
#pragma unroll
            for (int n = 0; n < 16; ++n) v[n] = p[pad_off(n * 16)];
            dftR<16>(v);
            twiddle_dfs<16>(v, wb2);
#pragma unroll
            for (int k = 0; k < 16; ++k) p[pad_off(k * 16)] = v[k];
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < WQA; ++e) load_window(e, wq[e]);
        // ---- pass 3 ----
        smem_pass<16 * NBT, T, 16, 1, true>(xch, nullptr, t, acc);
        cluster_arrive_relaxed();  // done with the exchange buffer: the next frame's first pass may overwrite it
    }
    cluster_wait();  // consume the last arrive; no peer writes to this CTA any more

    // ---- epilogue: this CTA's bins k = (c ROWS + r) + R0 k' are every CL-th group of ROWS output bins ----
    float* sout = reinterpret_cast<float*>(xch);  // N / CL floats
#pragma unroll
    for (int i = 0; i < NBT; ++i) {
        const int klow = PLN::low_freq(t + i * T + c * (ROWS * 256));
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int freq = klow + (N / 16) * jj;
            const int idx = (freq + N / 2) & (N - 1);
            const int li = (idx / R0) * ROWS + (idx % ROWS);
            sout[(ROWS == 4) ? (li ^ (((li >> 5) & 7) << 2)) : li] = acc[i * 16 + jj];
        }
    }
    __syncthreads();
    constexpr int NQL = N / CL / ROWS;  // groups of ROWS bins owned by this CTA
    if constexpr (ROWS == 4) {
        const float4* sout4 = reinterpret_cast<const float4*>(sout);
        for (int q = t; q < NQL; q += T) {
            float4 v = sout4[q ^ ((q >> 3) & 7)];
            const size_t qo = (size_t)q * CL + c;  // float4 index inside the column
            if (a.nsplit > 1) {
                reinterpret_cast<float4*>(a.partial + ((size_t)cs * a.nsplit + split) * N)[qo] = v;
            } else {
                v.x *= a.scale; v.y *= a.scale; v.z *= a.scale; v.w *= a.scale;
                const size_t o = (size_t)cs * (N / 4) + qo;
                if (a.out_lin) reinterpret_cast<float4*>(a.out_lin)[o] = v;
                if (a.out_db)
                    reinterpret_cast<float4*>(a.out_db)[o] = make_float4(power_to_db(v.x, a.eps), power_to_db(v.y, a.eps),
                                                                         power_to_db(v.z, a.eps), power_to_db(v.w, a.eps));
            }
        }
    } else {
        const float2* sout2 = reinterpret_cast<const float2*>(sout);
        for (int q = t; q < NQL; q += T) {
            float2 v = sout2[q];
            const size_t qo = (size_t)q * CL + c;  // float2 index inside the column
            if (a.nsplit > 1) {
                reinterpret_cast<float2*>(a.partial + ((size_t)cs * a.nsplit + split) * N)[qo] = v;
            } else {
                v.x *= a.scale; v.y *= a.scale;
                const size_t o = (size_t)cs * (N / 2) + qo;
                if (a.out_lin) reinterpret_cast<float2*>(a.out_lin)[o] = v;
                if (a.out_db) reinterpret_cast<float2*>(a.out_db)[o] = make_float2(power_to_db(v.x, a.eps), power_to_db(v.y, a.eps));
            }
        }
    }
}
