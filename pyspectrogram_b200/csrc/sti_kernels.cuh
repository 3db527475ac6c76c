// Fused STFT -> PSD -> STI kernels for sm_100a (device code).
//
// Work decomposition
//   A CTA owns one work item = (block of CPC (sub-channel, column) pairs, one chunk of those
//   columns' frames).  Its threads form F independent "frame groups" of T = N/E threads; a group
//   transforms one nfft-point frame at a time.  GPC groups cooperate on one column (frames
//   k0 + j*GPC + lane), so F = CPC * GPC.  Every pass is CTA-uniform: groups whose frame does not
//   exist transform zeros.
//
// Per frame
//   loader   TMA: one elected thread issues a cp.async.bulk (UBLKCP) per frame into a ring of
//            shared-memory stages, completion on an mbarrier; frames are contiguous complex64, so a
//            1-D bulk copy is exactly one frame.  The copy starts at the frame's address rounded
//            down to 16 B (frame starts come from np.linspace and are odd half the time), the
//            consumer skips the leading element.  LDG: strided/unaligned layouts load straight to
//            registers with coalesced LDG.64 (L1 no-allocate).
//   pass 0   registers <- window * samples; R0-point DFTs (radix <= 16, packed f32x2 math);
//            twiddle; store to the padded exchange buffer.
//   pass p   in-place R_p-point DFTs out of the exchange buffer; the last pass accumulates |X|^2
//            per bin in registers instead of storing.
//   epilogue (once per item) lanes of a column are summed, bins un-digit-reversed and fftshifted
//            through shared memory, scaled, and stored coalesced as linear power and/or
//            10*log10(p+eps) -- or as raw partial sums when a column is split over several CTAs
//            (summed in fp64 by sti_finalize_kernel, fixed order, deterministic).
//   Samples are read from HBM exactly once; nothing but the STI column is written.
//
// Index algebra (restated in numpy and checked by tests/test_fft_plan.py)
//   N = R0*R1*...*R(P-1),  S_p = N/(R0..Rp).  Before pass p the element with digits
//   (k_0..k_{p-1}, n_rest) sits at pos = sum_q k_q*S_q + n_rest.  Pass p splits
//   n_rest = n_p*S_p + n', does the R_p-point DFT over n_p, multiplies output k_p by
//   W_{R_p*S_p}^{n'*k_p} (skipped on the last pass) and stores it in place of n_p.  After the
//   last pass pos = sum_q k_q*S_q holds frequency k = k_0 + R0*k_1 + R0*R1*k_2 + ...
//   Shared-memory address of pos is pos + 2*(pos >> 4) (two complex of padding per 16): 64-bit
//   accesses of every pass stay conflict-free for the default plans, and a thread's run of
//   consecutive positions in the last pass is 16-byte aligned, so it is read with LDS.128 (the
//   LSU issues one warp instruction per 2 clocks whatever the width -- measured, tools/ubench).
//   Twiddles W_{R_p*S_p}^{n'*k} come from per-pass tables in one of two layouts:
//     column  twp[(k-1)*S_p + n']      lane-contiguous in n' (pass 0: loaded once into registers;
//                                      mid passes with S_p > 32: coalesced LDG.64 per frame)
//     row     twp[n'*(R_p+2) + k]      (mid passes with S_p <= 32) copied to shared memory once
//                                      per CTA, a thread's R_p-1 twiddles are one LDS.128 run.
#pragma once
#include <stdint.h>
#include "sti_common.cuh"

template <int N, int R0, int R1, int R2, int R3, int TWP = 0>
struct Plan {
    static constexpr int P = 1 + (R1 > 1) + (R2 > 1) + (R3 > 1);
    static constexpr int S0 = N / R0, S1 = S0 / R1, S2 = S1 / R2, S3 = S2 / R3;
    static_assert(R0 * R1 * R2 * R3 == N, "radices must multiply to N");
    static constexpr int RL = (P == 1) ? R0 : (P == 2) ? R1 : (P == 3) ? R2 : R3;  // last radix
    // offsets of the per-pass twiddle tables inside twp
    static constexpr bool ROW1 = (P >= 3) && (S1 <= 32), ROW2 = (P >= 4) && (S2 <= 32);
    static constexpr int TW0 = 0;
    static constexpr int TW1 = TW0 + (R0 - 1) * S0;
    static constexpr int ROWL1 = TWP ? 6 : R1 + 2, ROWL2 = TWP ? 6 : R2 + 2;  // complex per table row
    static constexpr int TW1_LEN = (P >= 3) ? (ROW1 ? S1 * ROWL1 : (TWP ? psg_npow(R1) : R1 - 1) * S1) : 0;
    static constexpr int TW2 = TW1 + TW1_LEN;
    static constexpr int TW2_LEN = (P >= 4) ? (ROW2 ? S2 * ROWL2 : (TWP ? psg_npow(R2) : R2 - 1) * S2) : 0;
    static constexpr int TW_TOTAL = TW2 + TW2_LEN;
    // frequency of (last-pass butterfly b, output j): digit-reverse b, add (N/RL)*j
    PSG_DEV static int low_freq(int b) {
        int rem = b, k = 0;
        if constexpr (P >= 2) { constexpr int s = S0 / RL; k += (rem / s); rem %= s; }
        if constexpr (P >= 3) { constexpr int s = S1 / RL; k += (rem / s) * R0; rem %= s; }
        if constexpr (P >= 4) { constexpr int s = S2 / RL; k += (rem / s) * R0 * R1; rem %= s; }
        return k;
    }
};

// psg_pad(base + n*S) == psg_pad(base) + pad_off(n*S) for every (base, S) the passes generate
// (n*S is either a multiple of 16 or added to a base whose low 4 bits leave room for it), so the
// per-element part of every shared-memory address is a compile-time immediate.
__host__ __device__ constexpr int pad_off(int d) { return d + 2 * (d >> 4); }

// twiddles of one mid pass for this thread's butterflies -> registers (issued before the barrier
// that precedes the pass, so their latency overlaps the wait).  ROW: from the shared-memory row
// table with LDS.128; else coalesced LDG.64 from the column table in global memory (L1 hits).
template <int E, int T, int R, int S, bool ROW, int TWP>
PSG_DEV void load_pass_tw(const float2* __restrict__ tab, int t, cf* tw, const cf* wbase = nullptr) {
    constexpr int NB = E / R;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int npr = (t + i * T) & (S - 1);
        if constexpr (TWP) {
            constexpr int NPW = psg_npow(R);
            cf pw[NPW];
            if constexpr (TWP == 2) {
                // W^1, W^2, W^4, W^8 of this butterfly are loop-invariant registers (correctly rounded
                // table values: squaring W^1 in fp32 would cost ~15 ulp on W^8)
#pragma unroll
                for (int q = 0; q < NPW; ++q) pw[q] = wbase[i * NPW + q];
            } else if constexpr (ROW) {
                const float4* row = reinterpret_cast<const float4*>(tab + npr * 6);
                const float4 v0 = row[0];
                pw[0] = make_float2(v0.x, v0.y);
                if constexpr (NPW >= 2) pw[1] = make_float2(v0.z, v0.w);
                if constexpr (NPW >= 3) {
                    const float4 v1 = row[1];
                    pw[2] = make_float2(v1.x, v1.y);
                    if constexpr (NPW >= 4) pw[3] = make_float2(v1.z, v1.w);
                }
            } else {
#pragma unroll
                for (int q = 0; q < NPW; ++q) pw[q] = __ldg(tab + q * S + npr);
            }
#pragma unroll
            for (int k = 1; k < R; ++k) {
                const int q = (k >= 8) ? 3 : (k >= 4) ? 2 : (k >= 2) ? 1 : 0;
                const int hb = 1 << q;
                tw[i * (R - 1) + k - 1] = (k == hb) ? pw[q] : cmul(tw[i * (R - 1) + (k - hb) - 1], pw[q]);
            }
        } else if constexpr (ROW) {
            const float4* row = reinterpret_cast<const float4*>(tab + npr * (R + 2));
            cf tmp[R];
#pragma unroll
            for (int k = 0; k < R / 2; ++k) {
                const float4 v = row[k];
                tmp[2 * k] = make_float2(v.x, v.y);
                tmp[2 * k + 1] = make_float2(v.z, v.w);
            }
#pragma unroll
            for (int k = 1; k < R; ++k) tw[i * (R - 1) + k - 1] = tmp[k];
        } else {
#pragma unroll
            for (int k = 1; k < R; ++k) tw[i * (R - 1) + k - 1] = __ldg(tab + (k - 1) * S + npr);
        }
    }
}

// one in-place pass over the exchange buffer (p >= 1). LAST: accumulate |X|^2 instead of storing.
template <int E, int T, int R, int S, bool LAST>
PSG_DEV void smem_pass(float2* __restrict__ buf, const cf* tw, int t, float* acc) {
    constexpr int NB = E / R;  // butterflies per thread
    constexpr int M = R * S;   // size of the sub-DFT this pass splits
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int b = t + i * T;
        const int npr = b & (S - 1);
        float2* p = buf + psg_pad((b / S) * M + npr);
        cf a[R];
        if constexpr (LAST && S == 1 && R >= 2) {
            // R consecutive positions inside one padded 16-block: 16-byte aligned, read as float4
            const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
            for (int n = 0; n < R / 2; ++n) {
                const float4 v = p4[n];
                a[2 * n] = make_float2(v.x, v.y);
                a[2 * n + 1] = make_float2(v.z, v.w);
            }
        } else {
#pragma unroll
            for (int n = 0; n < R; ++n) a[n] = p[pad_off(n * S)];
        }
        dftR<R>(a);
        if constexpr (LAST) {
#pragma unroll
            for (int k = 0; k < R; ++k) acc[i * R + k] = fmaf(a[k].x, a[k].x, fmaf(a[k].y, a[k].y, acc[i * R + k]));
        } else {
#pragma unroll
            for (int k = 1; k < R; ++k) a[k] = cmul(a[k], tw[i * (R - 1) + k - 1]);
#pragma unroll
            for (int k = 0; k < R; ++k) p[pad_off(k * S)] = a[k];
        }
    }
}

// Pass p (radix 16, stride 32) with the following radix-2 pass (stride 16) done in registers: the two
// inputs of a radix-2 butterfly sit in lanes l and l ^ 16 of one warp (same k, n' and n' + 16), so each
// lane sends half of its 16 outputs to its partner with SHFL (8 complex = 16 SHFL.32, half the LSU
// cost of re-reading them) and finishes the butterflies of the other half: lanes 0..15 those of
// k = 0..7, lanes 16..31 those of k = 8..15.  The pair reads and writes the same 32 positions, so the
// pass stays in place, and a warp writes exactly the 512 positions its own threads read in the last
// (radix-16, stride 1) pass.  One shared-memory write + read less than a four-pass plan
// (8192 = 16*16*2*16: 6.5 instead of 8 accesses per sample).  w2 = W_32^(lane & 15), negated in the
// upper half-warp (there the difference is taken the other way round).
// the exchange itself: a[k] = this lane's 16 outputs (n' = lane), q = padded address of position
// (block base + (n' & 15)); stores the finished radix-2 outputs of the eight k this lane keeps
PSG_DEV void pair_exchange_store(const cf* a, float2* q, const bool hi, const cf w2) {
    if (hi) q += pad_off(8 * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const cf send = hi ? a[i] : a[8 + i];
        const cf keep = hi ? a[8 + i] : a[i];
        cf recv;
        recv.x = __shfl_xor_sync(0xffffffffu, send.x, 16);
        recv.y = __shfl_xor_sync(0xffffffffu, send.y, 16);
        q[pad_off(i * 32)] = cadd(keep, recv);
        q[pad_off(i * 32) + 18] = cmul(csub(keep, recv), w2);
    }
}
template <int E, int T, int R, int S>
PSG_DEV void smem_pass_fused_r2(float2* __restrict__ buf, const cf* tw, int t, const cf w2) {
    static_assert(E == 16 && R == 16 && S == 32 && (T % 32) == 0, "pair exchange is laid out for radix 16, stride 32");
    const int b = t;
    const int npr = b & 31;
    const bool hi = (npr & 16) != 0;
    float2* p = buf + psg_pad((b / S) * (R * S) + npr);
    cf a[R];
#pragma unroll
    for (int n = 0; n < R; ++n) a[n] = p[pad_off(n * S)];
    dftR<R>(a);
#pragma unroll
    for (int k = 1; k < R; ++k) a[k] = cmul(a[k], tw[k - 1]);
    pair_exchange_store(a, buf + psg_pad((b / S) * (R * S) + (npr & 15)), hi, w2);
}

// The exchange between pass p (radix Ra, stride Sa) and pass p+1 (radix Rb) stays inside aligned
// groups of Sa consecutive threads when both passes run one butterfly per thread (Ra == Rb == E):
// pass p's threads [q*Sa, (q+1)*Sa) write exactly the blocks those same threads read in pass p+1.
// With Sa <= 32 that group lives in one warp and __syncwarp() replaces the CTA barrier.
template <int E, int Ra, int Rb, int Sa>
struct WarpLocal {
    static constexpr bool value = (Ra == E) && (Rb == E) && (Sa <= 32);
};
template <bool LOCAL>
PSG_DEV void exchange_sync() {
    if constexpr (LOCAL) __syncwarp();
    else __syncthreads();
}

template <int LOGN, int E, int R0, int R1, int R2, int R3, int F, int LOADER, int STAGES, int XBUF, int IQT = IQ_C64,
          int TWP = 0>
struct FusedCfg {
    static constexpr int N = 1 << LOGN, T = N / E, NT = F * T;
    using PLN = Plan<N, R0, R1, R2, R3, TWP>;
    static constexpr int NPAD = psg_pad(N) + 2;  // even: every group's buffer stays 16-byte aligned
    static constexpr int SLOT = N * IqBytes<IQT>::value + 16;  // bytes: one frame + 16 of alignment slack
    static constexpr int TWSM = (TWP == 2) ? 0 : (PLN::ROW1 ? PLN::TW1_LEN : 0) + (PLN::ROW2 ? PLN::TW2_LEN : 0);  // complex
    static constexpr size_t bar_bytes = 64 + (size_t)TWSM * 8;
    static constexpr size_t stage_bytes = (LOADER == PSG_LOADER_TMA) ? (size_t)STAGES * F * SLOT : 0;
    static constexpr size_t xch_bytes = (size_t)F * XBUF * NPAD * 8;
    static constexpr size_t smem_bytes = bar_bytes + stage_bytes + xch_bytes;
};

// MULTI (one-frame columns, Mode R): the CTA runs a.cb consecutive column blocks back to back, one
// frame per group and iteration, with the epilogue inside the loop; the TMA producer looks ahead
// across columns, so table loads, barrier setup and the load latency are paid once per CTA instead
// of once per frame.
template <int LOGN, int E, int R0, int R1, int R2, int R3, int F, int LOADER, int STAGES, int XBUF, int MINB, int PFL2 = 0,
          int IQT = IQ_C64, int TWP = 0, int MULTI = 0>
__global__ void __launch_bounds__(F*((1 << LOGN) / E), MINB) sti_fused_kernel(const StiArgs a) {
    using CF = FusedCfg<LOGN, E, R0, R1, R2, R3, F, LOADER, STAGES, XBUF, IQT, TWP>;
    constexpr int IQB = IqBytes<IQT>::value;
    constexpr int N = CF::N, T = CF::T, NT = CF::NT, NPAD = CF::NPAD, SLOT = CF::SLOT;
    using PL = Plan<N, R0, R1, R2, R3, TWP>;
    constexpr int P = PL::P;
    constexpr int NB0 = E / R0;
    static_assert(P >= 2, "tuned kernels need at least two passes");
    static_assert(R0 <= E && R1 <= E && R2 <= E && R3 <= E, "a thread must hold a whole butterfly");
    static_assert(STAGES <= 8, "one mbarrier per stage in a 64-byte header");
    constexpr bool L01 = WarpLocal<E, R0, R1, PL::S0>::value;
    constexpr bool L12 = WarpLocal<E, R1, R2, PL::S1>::value;
    constexpr bool L23 = WarpLocal<E, R2, R3, PL::S2>::value;
    // 16 x 16 x 2 x 16: the radix-2 pass runs in registers on the end of pass 1 (smem_pass_fused_r2)
    constexpr bool FUSE2 = (P == 4) && (E == 16) && (R1 == 16) && (R2 == 2) && (R3 == 16) && (PL::S1 == 32);
    // 16 x 2 x 16 (512 points, one warp per frame): the same exchange on the end of pass 0 -- one
    // shared-memory round trip per frame, no CTA barrier
    constexpr bool FUSE0 = (P == 3) && (E == 16) && (R0 == 16) && (R1 == 2) && (R2 == 16) && (PL::S0 == 32) && (T == 32);
    // true when no exchange needs a CTA barrier: every frame group then owns its buffers privately
    constexpr bool ALL_LOCAL = FUSE0 || (L01 && (P < 3 || L12) && (P < 4 || L23) && !FUSE2);
    static_assert(ALL_LOCAL ? (T <= 32) : true, "");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    float2* twsm = reinterpret_cast<float2*>(smem_raw + 64);  // row-layout twiddle tables
    unsigned char* stage = smem_raw + CF::bar_bytes;
    float2* xch = reinterpret_cast<float2*>(smem_raw + CF::bar_bytes + CF::stage_bytes);

    const int tid = threadIdx.x;
    const int g = tid / T;   // frame group
    const int t = tid - g * T;
    const int gpc = a.gpc;   // groups per column
    const int cpc = F / gpc; // columns per CTA
    const int lane = g % gpc;
    const int slot = g / gpc;

    const int ncs = a.ncol * a.nsub;
    const int item = blockIdx.x;
    const int split = MULTI ? 0 : item % a.nsplit;
    // MULTI: cs0 is the first column of the item's first block; iteration j works on block j
    const int cs0 = MULTI ? item * a.cb * cpc : (item / a.nsplit) * cpc;
    const int cs = cs0 + slot;
    const bool col_ok = cs < ncs;
    const int k0 = split * a.chunk;
    const int k1 = min(a.nfr, k0 + a.chunk);
    const int niter = MULTI ? min(a.cb, (ncs + cpc - 1) / cpc - item * a.cb) : (k1 - k0 + gpc - 1) / gpc;

    long long fbase = 0;  // element offset of frame k0+lane of this group's column
    if (!MULTI && col_ok) {
        const int col = cs % a.ncol, sub = cs / a.ncol;
        fbase = a.col_off[col] + (long long)sub * a.sub_stride + (long long)(k0 + lane) * a.hop_elems;
    }
    const long long fstep = (long long)gpc * a.hop_elems;
    // MULTI: the one frame of this group in iteration j (column block j of the item)
    auto frame_of = [&](int j, bool& ok) -> long long {
        const int c = cs0 + j * cpc + slot;
        ok = c < ncs && lane < a.nfr;
        if (!ok) return 0;
        const int col = c % a.ncol, sub = c / a.ncol;
        return a.col_off[col] + (long long)sub * a.sub_stride + (long long)lane * a.hop_elems;
    };

    // ---- TMA producer: thread 0 of every frame group fetches its own group's frames ----
    // iteration pj -> stage pj % STAGES, slot g; every group leader arrives once per phase
    long long pbase = fbase;  // producer cursor: frame of the next iteration to fetch
    int pj = 0;               // next iteration to fetch
    auto produce = [&]() {
        if constexpr (LOADER == PSG_LOADER_TMA) {
            uint64_t* bar = bars + (pj % STAGES);
            bool pok = col_ok && (k0 + pj * gpc + lane) < k1;
            if constexpr (MULTI) pbase = frame_of(pj, pok);
            if (pok) {
                const uintptr_t src = reinterpret_cast<uintptr_t>(a.iq) + (uintptr_t)(pbase * IQB);
                const uint32_t bytes = N * IQB + ((src & 15) ? 16 : 0);
                mbar_expect_tx(bar, bytes);
                bulk_g2s(stage + ((size_t)(pj % STAGES) * F + g) * SLOT, reinterpret_cast<const void*>(src & ~(uintptr_t)15),
                         bytes, bar);
            } else {
                mbar_expect_tx(bar, 0);
            }
            ++pj;
            pbase += fstep;
        }
    };
    if constexpr (CF::TWSM > 0) {
        const float2* src = a.twp + (PL::ROW1 ? PL::TW1 : PL::TW2);
        for (int i = tid; i < CF::TWSM; i += NT) twsm[i] = __ldg(src + i);
    }
    if constexpr (LOADER == PSG_LOADER_TMA) {
        if (tid == 0) {
            for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, F);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if constexpr (LOADER == PSG_LOADER_TMA || CF::TWSM > 0) __syncthreads();
    if constexpr (LOADER == PSG_LOADER_TMA) {
        // prefetch distance: STAGES-1 iterations; a single stage is refilled right after its reads
        if (t == 0)
            for (int jj = 0; jj < (STAGES == 1 ? 1 : STAGES - 1) && jj < niter; ++jj) produce();
    }

    // loop-invariant tables in registers: window of this thread's E samples, pass-0 twiddles
    float w[E];
    cf tw0[NB0 * (R0 - 1)];
#pragma unroll
    for (int i = 0; i < NB0; ++i) {
        const int b = t + i * T;
#pragma unroll
        for (int n = 0; n < R0; ++n) w[i * R0 + n] = __ldg(a.win + b + n * PL::S0);
#pragma unroll
        for (int k = 1; k < R0; ++k) tw0[i * (R0 - 1) + k - 1] = __ldg(a.twp + PL::TW0 + (k - 1) * PL::S0 + b);
    }
    // TWP == 2: the power-of-two twiddles of every mid-pass butterfly of this thread stay in registers
    constexpr int NPW1 = psg_npow(R1), NPW2 = psg_npow(R2);
    cf wb1[(TWP == 2 && P >= 3 && !FUSE0) ? (E / R1) * NPW1 : 1], wb2[(TWP == 2 && P >= 4 && !FUSE2) ? (E / R2) * NPW2 : 1];
    if constexpr (TWP == 2 && P >= 3 && !FUSE0) {
#pragma unroll
        for (int i = 0; i < E / R1; ++i) {
            const int npr = (t + i * T) & (PL::S1 - 1);
#pragma unroll
            for (int q = 0; q < NPW1; ++q) wb1[i * NPW1 + q] = __ldg(a.twp + PL::TW1 + (PL::ROW1 ? npr * 6 + q : q * PL::S1 + npr));
        }
    }
    cf w2f = make_float2(1.f, 0.f);  // FUSE2: W_32^(lane & 15), sign of the upper half-warp folded in
    if constexpr (FUSE2 || FUSE0) {
        static_assert(!(FUSE2 || FUSE0) || TWP == 2, "fused radix-2 plans use the power-layout tables");
        w2f = __ldg(a.twp + (FUSE2 ? PL::TW2 : PL::TW1) + (t & 15) * 6);
        if (t & 16) w2f = make_float2(-w2f.x, -w2f.y);
    }
    if constexpr (TWP == 2 && P >= 4 && !FUSE2) {
#pragma unroll
        for (int i = 0; i < E / R2; ++i) {
            const int npr = (t + i * T) & (PL::S2 - 1);
#pragma unroll
            for (int q = 0; q < NPW2; ++q) wb2[i * NPW2 + q] = __ldg(a.twp + PL::TW2 + (PL::ROW2 ? npr * 6 + q : q * PL::S2 + npr));
        }
    }
    float acc[E];  // |X|^2 sums of this thread's E bins (scalar: registers are the scarce resource)
#pragma unroll
    for (int i = 0; i < E; ++i) acc[i] = 0.f;

    // ---- epilogue: digit-reversed register sums -> fftshifted, coalesced 128-bit stores ----
    // (one frame per column in Mode R makes this as hot as the transform: keep it lean)
    auto epilogue = [&](const int cbase, float* acc) {
        __syncthreads();
        float* sout = reinterpret_cast<float*>(xch);  // [F][N] floats (fits: F*NPAD*8 bytes available)
        constexpr int RL = PL::RL;
        constexpr int NBL = E / RL;
        // bins are swizzled in groups of four (bits 2..4 ^= bits 5..7) so that the scattered 32-bit stores
        // spread over the banks while every aligned group of four bins stays one 128-bit word
    #pragma unroll
        for (int i = 0; i < NBL; ++i) {
            const int b = t + i * T;
            const int klow = PL::low_freq(b);
    #pragma unroll
            for (int jj = 0; jj < RL; ++jj) {
                const int freq = klow + (N / RL) * jj;
                const int idx = (freq + N / 2) & (N - 1);
                sout[g * N + (idx ^ (((idx >> 5) & 7) << 2))] = acc[i * RL + jj];
            }
        }
        __syncthreads();
        // all NT threads cooperate: column slot s, four bins at idx; lanes summed in fixed order
        constexpr int NQ = N / 4;
        const float4* sout4 = reinterpret_cast<const float4*>(sout);
        for (int e = tid; e < cpc * NQ; e += NT) {
            const int s = e / NQ, q = e - s * NQ;
            const int c = cbase + s;
            if (c >= ncs) break;
            const int sw = q ^ ((q >> 3) & 7);  // the swizzle above, in units of four bins
            float4 v = sout4[(s * gpc) * NQ + sw];
            for (int l = 1; l < gpc; ++l) {
                const float4 u = sout4[(s * gpc + l) * NQ + sw];
                v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
            }
            if (a.nsplit > 1) {
                reinterpret_cast<float4*>(a.partial + ((size_t)c * a.nsplit + split) * N)[q] = v;
            } else {
                v.x *= a.scale; v.y *= a.scale; v.z *= a.scale; v.w *= a.scale;
                const size_t o = (size_t)c * NQ + q;
                if (a.out_lin) reinterpret_cast<float4*>(a.out_lin)[o] = v;
                if (a.out_db)
                    reinterpret_cast<float4*>(a.out_db)[o] = make_float4(power_to_db(v.x, a.eps), power_to_db(v.y, a.eps),
                                                                         power_to_db(v.z, a.eps), power_to_db(v.w, a.eps));
            }
        }
    };
    for (int j = 0; j < niter; ++j, fbase += fstep) {
        bool valid = col_ok && (k0 + j * gpc + lane) < k1;
        if constexpr (MULTI) fbase = frame_of(j, valid);
        float2* buf = xch + (size_t)(g * XBUF + (XBUF > 1 ? (j & 1) : 0)) * NPAD;
        cf x[E];
        // ---- pass 0: samples -> registers ----
        if constexpr (LOADER == PSG_LOADER_TMA) {
            // slot g of stage (j-1) % STAGES was last read by this group in iteration j-1; all its
            // threads are past those reads (CTA barrier of iteration j-1, or the group is one warp)
            if constexpr (ALL_LOCAL) __syncwarp();
            if (STAGES > 1 && t == 0 && pj < niter) produce();
            mbar_wait(bars + (j % STAGES), (j / STAGES) & 1);
            const unsigned char* sb = stage + ((size_t)(j % STAGES) * F + g) * SLOT;
            // the copy started at the 16-byte boundary below the frame: skip the leading elements
            const int skew = (int)(((reinterpret_cast<uintptr_t>(a.iq) + (uintptr_t)(fbase * IQB)) & 15) / IQB);
            if (valid) {
#pragma unroll
                for (int i = 0; i < NB0; ++i)
#pragma unroll
                    for (int n = 0; n < R0; ++n) x[i * R0 + n] = lds_iq<IQT>(sb, skew + t + i * T + n * PL::S0);
            } else {
#pragma unroll
                for (int i = 0; i < E; ++i) x[i] = make_float2(0.f, 0.f);
            }
        } else {
            if constexpr (PFL2 > 0) {
                // contiguous frames: pull the frame PFL2 iterations ahead into L2 so that the LDGs of
                // that iteration see L2 latency, not DRAM latency (no registers, no shared memory)
                if (t == 0 && a.sample_stride == 1 && col_ok && (k0 + (j + PFL2) * gpc + lane) < k1) {
                    const uintptr_t pa = (reinterpret_cast<uintptr_t>(a.iq) + (uintptr_t)((fbase + (long long)PFL2 * fstep) * IQB)) &
                                         ~(uintptr_t)15;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pa), "r"(N * IQB + 16) : "memory");
                }
            }
            if (valid) {
#pragma unroll
                for (int i = 0; i < NB0; ++i)
#pragma unroll
                    for (int n = 0; n < R0; ++n)
                        x[i * R0 + n] = ldg_iq<IQT>(a.iq, fbase + (long long)(t + i * T + n * PL::S0) * a.sample_stride);
            } else {
#pragma unroll
                for (int i = 0; i < E; ++i) x[i] = make_float2(0.f, 0.f);
            }
        }
        // previous frame's last pass must be done reading before this buffer is overwritten
        if constexpr (XBUF == 1) exchange_sync<ALL_LOCAL>();
#pragma unroll
        for (int i = 0; i < NB0; ++i) {
            dftRw<R0>(&x[i * R0], &w[i * R0]);  // window multiply folded into the first butterfly layer
            float2* p0 = buf + psg_pad(t + i * T);
#pragma unroll
            for (int kk = 1; kk < R0; ++kk) x[i * R0 + kk] = cmul(x[i * R0 + kk], tw0[i * (R0 - 1) + kk - 1]);
            if constexpr (FUSE0) {
                pair_exchange_store(&x[0], buf + psg_pad(t & 15), (t & 16) != 0, w2f);
            } else {
#pragma unroll
                for (int kk = 0; kk < R0; ++kk) p0[pad_off(kk * PL::S0)] = x[i * R0 + kk];
            }
        }
        if constexpr (FUSE0) {
            __syncwarp();
            if constexpr (LOADER == PSG_LOADER_TMA && STAGES == 1) {
                if (t == 0 && pj < niter) produce();  // the group's slot has been read by its one warp
            }
            smem_pass<E, T, R2, PL::S2, true>(buf, nullptr, t, acc);
        } else {
            cf tw1[(P >= 3) ? (E / R1) * (R1 - 1) : 1];
            if constexpr (P >= 3) load_pass_tw<E, T, R1, PL::S1, PL::ROW1, TWP>(PL::ROW1 ? twsm : a.twp + PL::TW1, t, tw1, wb1);
            exchange_sync<L01>();
            if constexpr (LOADER == PSG_LOADER_TMA && STAGES == 1) {
                static_assert(LOADER != PSG_LOADER_TMA || STAGES > 1 || !L01, "single stage needs a CTA barrier");
                if (t == 0 && pj < niter) produce();  // the stage has been read by everyone: refill it
            }
            if constexpr (FUSE2) smem_pass_fused_r2<E, T, R1, PL::S1>(buf, tw1, t, w2f);
            else smem_pass<E, T, R1, PL::S1, P == 2>(buf, tw1, t, acc);
        }
        if constexpr (P >= 3 && !FUSE2 && !FUSE0) {
            cf tw2[(P >= 4) ? (E / R2) * (R2 - 1) : 1];
            if constexpr (P >= 4)
                load_pass_tw<E, T, R2, PL::S2, PL::ROW2, TWP>(PL::ROW2 ? twsm + (PL::ROW1 ? PL::TW1_LEN : 0) : a.twp + PL::TW2, t, tw2, wb2);
            exchange_sync<L12>();
            smem_pass<E, T, R2, PL::S2, P == 3>(buf, tw2, t, acc);
        }
        if constexpr (P >= 4) {
            exchange_sync<L23 || FUSE2>();  // FUSE2: a warp reads back what it wrote (see smem_pass_fused_r2)
            smem_pass<E, T, R3, PL::S3, true>(buf, nullptr, t, acc);
        }
        if constexpr (MULTI) {  // every frame is a finished column
            epilogue(cs0 + j * cpc, acc);
#pragma unroll
            for (int i = 0; i < E; ++i) acc[i] = 0.f;
            __syncthreads();  // the staging area of the epilogue aliases the exchange buffers
        }
    }
    if constexpr (!MULTI) epilogue(cs0, acc);


}

// Sum the partial columns of split work items (fixed order -> deterministic), scale, store.
__global__ void sti_finalize_kernel(const float* __restrict__ partial, int nsplit, int n, size_t ncols_total,
                                    float scale, float eps, float* out_lin, float* out_db) {
    const size_t total = ncols_total * (size_t)n;
    for (size_t gi = blockIdx.x * (size_t)blockDim.x + threadIdx.x; gi < total;
         gi += (size_t)gridDim.x * blockDim.x) {
        const size_t c = gi / n;
        const int i = (int)(gi - c * n);
        const float* p = partial + c * nsplit * (size_t)n + i;
        double s = 0.0;
        for (int k = 0; k < nsplit; ++k) s += (double)p[(size_t)k * n];
        const float v = (float)(s * (double)scale);
        if (out_lin) out_lin[gi] = v;
        if (out_db) out_db[gi] = power_to_db(v, eps);
    }
}

// ---- generic kernels: any power-of-two N, radix-2, simple ---------------------------------------
// Used for N < 256, for N the tuned kernels do not cover, and to cross-check them; never for speed.
// work buffer: shared memory (N <= 16384) or a per-CTA global scratch slab (larger N).
__global__ void __launch_bounds__(256) sti_generic_kernel(const StiArgs a, int logn, float2* gwork, float* gacc) {
    const int N = 1 << logn;
    extern __shared__ __align__(16) float2 smem[];
    float2* buf = gwork ? gwork + (size_t)blockIdx.x * N : smem;
    float* accs = gacc ? gacc + (size_t)blockIdx.x * N : reinterpret_cast<float*>(smem + N);
    const int t = threadIdx.x, nt = blockDim.x;
    const int ncs = a.ncol * a.nsub;
    for (int item = blockIdx.x; item < ncs * a.nsplit; item += gridDim.x) {
        const int split = item % a.nsplit;
        const int cs = item / a.nsplit;
        const int col = cs % a.ncol;
        const int sub = cs / a.ncol;
        const int k0 = split * a.chunk;
        const int k1 = min(a.nfr, k0 + a.chunk);
        long long src = a.col_off[col] + (long long)sub * a.sub_stride + (long long)k0 * a.hop_elems;
        __syncthreads();
        for (int i = t; i < N; i += nt) accs[i] = 0.f;
        for (int k = k0; k < k1; ++k, src += a.hop_elems) {
            __syncthreads();
            for (int i = t; i < N; i += nt)
                buf[i] = cscale(ldg_iq_rt(a.iq_type, a.iq, src + (long long)i * a.sample_stride), a.win[i]);
            for (int s = N >> 1; s >= 1; s >>= 1) {
                __syncthreads();
                const int step = (N >> 1) / s;  // twiddle stride: W_{2s}^j = W_N^{j*N/(2s)}
                for (int i = t; i < (N >> 1); i += nt) {
                    const int j = i & (s - 1);
                    const int base = ((i / s) * 2 * s) + j;
                    const cf u = buf[base], v = buf[base + s];
                    buf[base] = cadd(u, v);
                    cf d = csub(u, v);
                    buf[base + s] = (s > 1) ? cmul(d, a.tw[(size_t)j * step]) : d;
                }
            }
            __syncthreads();
            for (int i = t; i < N; i += nt) {
                const cf v = buf[i];
                accs[i] = fmaf(v.x, v.x, fmaf(v.y, v.y, accs[i]));
            }
        }
        __syncthreads();
        for (int i = t; i < N; i += nt) {
            const int freq = (int)(__brev((unsigned)i) >> (32 - logn));
            const int idx = (freq + N / 2) & (N - 1);
            if (a.nsplit > 1) {
                a.partial[((size_t)cs * a.nsplit + split) * N + idx] = accs[i];
            } else {
                const float p = accs[i] * a.scale;
                if (a.out_lin) a.out_lin[(size_t)cs * N + idx] = p;
                if (a.out_db) a.out_db[(size_t)cs * N + idx] = power_to_db(p, a.eps);
            }
        }
    }
}

// ---- median over the time axis (np.median(sxx, axis=1), drfProc.py:401 / :451) ------------------
// img [nsub][ncol][nfft]; exact k-th order statistic per (sub, bin) by a 32-step bisection on the
// IEEE bit pattern (monotone for the non-negative powers this path produces; negative values are
// mapped to an order-preserving key as well).
PSG_DEV unsigned f2key(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
PSG_DEV float key2f(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// One CTA owns BPC adjacent bins of one sub-channel and all ncol columns of them; its 256 threads
// are BPC bins x SL = 256/BPC column slices.  The [ncol][BPC] tile is read from global memory once
// (rows of BPC floats, coalesced), converted to order-preserving keys and kept in shared memory
// (TILED; ncol*BPC*4 bytes) -- or, when the tile does not fit, re-read from L2 on every step.
// Each bisection step: every thread counts its slice, slice counts meet in shared memory, every
// thread of a bin adds them up in the same order, so all SL threads of a bin walk the same path.
template <int BPC, bool TILED>
__global__ void __launch_bounds__(256) median_time_kernel(const float* __restrict__ img, int nsub, int ncol,
                                                          int nfft, float eps, float* med_lin, float* med_db) {
    constexpr int SL = 256 / BPC;
    extern __shared__ __align__(16) unsigned med_smem[];
    unsigned* cnts = med_smem;                 // [2][SL][BPC] slice counts, ping-pong
    unsigned* mins = med_smem + 2 * 256;       // [SL][BPC]
    unsigned* tile = med_smem + 3 * 256;       // [ncol][BPC] keys (TILED)
    const int tid = threadIdx.x;
    const int b = tid % BPC, s = tid / BPC;
    const int blocks_per_sub = (nfft + BPC - 1) / BPC;
    const int sub = blockIdx.x / blocks_per_sub;
    const int bin0 = (blockIdx.x - sub * blocks_per_sub) * BPC;
    const int bin = bin0 + b;
    const bool live = bin < nfft;
    const float* p = img + (size_t)sub * ncol * nfft + (live ? bin : nfft - 1);
    if constexpr (TILED) {
        for (int c = s; c < ncol; c += SL) tile[c * BPC + b] = f2key(__ldg(p + (size_t)c * nfft));
        __syncthreads();
    }
    auto key_at = [&](int c) -> unsigned {
        if constexpr (TILED) return tile[c * BPC + b];
        else return f2key(__ldg(p + (size_t)c * nfft));
    };
    const int klo = (ncol - 1) >> 1;  // 0-based rank of the lower middle
    // largest key K with count(key < K) <= klo  ==> K is the klo-th smallest key
    unsigned key = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const unsigned cand = key | (1u << bit);
        unsigned cnt = 0;
        for (int c = s; c < ncol; c += SL) cnt += (key_at(c) < cand) ? 1u : 0u;
        unsigned* cb = cnts + (bit & 1) * 256;
        cb[s * BPC + b] = cnt;
        __syncthreads();
        unsigned tot = 0;
#pragma unroll
        for (int q = 0; q < SL; ++q) tot += cb[q * BPC + b];
        if (tot <= (unsigned)klo) key = cand;
    }
    float m = key2f(key);
    if ((ncol & 1) == 0) {
        // upper middle: the same value if it is repeated, else the smallest value above it
        unsigned cnt_le = 0, nxt = 0xffffffffu;
        for (int c = s; c < ncol; c += SL) {
            const unsigned kk = key_at(c);
            cnt_le += (kk <= key) ? 1u : 0u;
            if (kk > key && kk < nxt) nxt = kk;
        }
        // buffer 1 was last read in step bit=1, and every thread is past the barrier of step bit=0
        cnts[256 + s * BPC + b] = cnt_le;
        mins[s * BPC + b] = nxt;
        __syncthreads();
        unsigned tot = 0, mn = 0xffffffffu;
#pragma unroll
        for (int q = 0; q < SL; ++q) {
            tot += cnts[256 + q * BPC + b];
            mn = min(mn, mins[q * BPC + b]);
        }
        const float hi = (tot >= (unsigned)klo + 2u) ? m : key2f(mn);
        m = (m + hi) * 0.5f;  // numpy: mean of the two middle values in float32
    }
    if (live && s == 0) {
        const size_t o = (size_t)sub * nfft + bin;
        if (med_lin) med_lin[o] = m;
        if (med_db) med_db[o] = power_to_db(m, eps);
    }
}

// ---- time-median, warp-per-bin selection (the default; median_time_kernel above is the fallback for column
// counts whose tile does not fit shared memory) -------------------------------------------------------------
// The bisection kernel above runs 32 CTA-wide steps -- a CTA barrier each, eight warps per SM at thousands of
// columns -- and reaches 4-5 % of the HBM peak: it is bound by latency, not by its ~96 instructions per key.
// Here a WARP owns a bin: its ncol keys sit bin-major in shared memory and nothing but the tile load
// synchronises the CTA.  The selection is the same exact bit-by-bit search for the key of rank (ncol-1)/2
// ("largest K with count(key < K) <= rank"), made cheaper four ways:
//   1. it starts at the highest bit in which the bin's keys differ (min ^ max; powers of one bin share sign and
//      most of the exponent), and a constant bin is answered at once;
//   2. after every LEVEL_BITS steps only the keys that match the decided prefix can still change a count: they are
//      compacted in place to the front of the row (ballot + popc; a write never passes the read position) and
//      the next steps run on those few per cent of the keys; keys below the prefix range are remembered as a
//      count, keys above it as their minimum.  (Measured alternative: every lane compacting its own strided keys
//      without ballots -- fewer instructions, 15 % slower: its scalar loads and divergent stores are latency-bound);
//   3. once at most 32 keys are left (two levels at 3600 columns) they are ranked against one another in
//      registers with shuffles -- the per-level overhead of ~500 warp instructions a bin is what the small sets
//      cost, not their keys;
//   4. even ncol: count(key <= K) and the next larger key -- numpy's mean of the two middle values in float32 --
//      come out of the same bookkeeping, no further pass over the row.
// Result: bit-identical to median_time_kernel and to np.median for the finite, non-NaN powers this path produces.
template <int BPC>
__global__ void __launch_bounds__(BPC * 32) median_select_kernel(const float* __restrict__ img, int nsub, int ncol, int nfft,
                                                                int rowlen, float eps, float* med_lin, float* med_db) {
    constexpr int LEVEL_BITS = 4;
    extern __shared__ __align__(16) unsigned msel_smem[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int blocks_per_sub = (nfft + BPC - 1) / BPC;
    const int sub = blockIdx.x / blocks_per_sub;
    const int bin0 = (blockIdx.x - sub * blocks_per_sub) * BPC;
    {
        // rows of BPC floats per column (coalesced per row; adjacent CTAs share sectors through L2), copied
        // asynchronously (LDGSTS, 4 bytes each) straight to their bin-major place: every load of the tile is in
        // flight at once and no register waits for it -- with a dozen resident warps per SM a register-staged
        // load spent a third of the kernel on DRAM latency.  Pads are 0xffffffff (never below a candidate).
        const int b = tid % BPC, s0 = tid / BPC;
        const int bin = min(bin0 + b, nfft - 1);
        const float* src = img + (size_t)sub * ncol * nfft + bin + (size_t)s0 * nfft;
        const size_t step = (size_t)32 * nfft;
        unsigned* row = msel_smem + (size_t)b * rowlen;
        uint32_t dst = smem_u32(row + s0);
        int c = s0;
#pragma unroll 4
        for (; c < ncol; c += 32, src += step, dst += 128)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
        for (; c < rowlen; c += 32) row[c] = 0xffffffffu;
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int bin = bin0 + w;
    if (bin >= nfft) return;
    unsigned* const row = msel_smem + (size_t)w * rowlen;
    const unsigned klo = (unsigned)((ncol - 1) >> 1);  // 0-based rank of the lower middle
    // raw floats -> order-preserving keys in place, minimum and maximum on the way
    unsigned mn = 0xffffffffu, mx = 0u;
    {
        uint4* const row4 = reinterpret_cast<uint4*>(row);
        const int full = ncol >> 2;  // groups of four that hold no pad
        for (int i = lane; i < full; i += 32) {
            uint4 k = row4[i];
            k.x = f2key(__uint_as_float(k.x));
            k.y = f2key(__uint_as_float(k.y));
            k.z = f2key(__uint_as_float(k.z));
            k.w = f2key(__uint_as_float(k.w));
            row4[i] = k;
            mn = min(mn, min(min(k.x, k.y), min(k.z, k.w)));
            mx = max(mx, max(max(k.x, k.y), max(k.z, k.w)));
        }
        const int i = 4 * full + lane;
        if (i < ncol) {
            const unsigned k = f2key(__uint_as_float(row[i]));
            row[i] = k;
            mn = min(mn, k);
            mx = max(mx, k);
        }
        __syncwarp();
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    unsigned key = mn, cnt_le = (unsigned)ncol, nxt = 0xffffffffu;  // a constant bin: every key is the answer
    if (mn != mx) {
        int bit = 31 - __clz((int)(mn ^ mx));                    // highest differing bit
        key = (bit == 31) ? 0u : (mx >> (bit + 1)) << (bit + 1);  // the bits every key shares
        unsigned below = 0;  // count(row < key)
        unsigned base = 0;   // keys below the range of the active ones (dropped at the last compaction)
        int na = ncol;       // active keys: row[0 .. na)
        bool done = false;
        const unsigned lt = (1u << lane) - 1u;
        while (bit >= 0) {
            if (na <= 32) {
                // one key per lane left: rank them against one another in registers (ties broken by lane) and pick
                // the one of rank klo - base: no more loops, per-bit reductions or compactions
                const unsigned kv = (lane < na) ? row[lane] : 0xffffffffu;
                unsigned rank = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const unsigned kj = __shfl_sync(0xffffffffu, kv, j);
                    rank += (kj < kv || (kj == kv && j < lane)) ? 1u : 0u;
                }
                const unsigned sel = __ballot_sync(0xffffffffu, lane < na && rank == klo - base);
                key = __shfl_sync(0xffffffffu, kv, __ffs((int)sel) - 1);
                cnt_le = base + __popc(__ballot_sync(0xffffffffu, lane < na && kv <= key));
                nxt = __reduce_min_sync(0xffffffffu, min(nxt, (lane < na && kv > key) ? kv : 0xffffffffu));
                done = true;
                break;
            }
            // up to LEVEL_BITS steps on the active keys (padded to a multiple of four with 0xffffffff)
            const int stop = max(bit - LEVEL_BITS + 1, 0);
            const int n4 = (na + 3) >> 2;
            if (lane < ((4 - (na & 3)) & 3)) row[na + lane] = 0xffffffffu;
            __syncwarp();
            const uint4* const row4 = reinterpret_cast<const uint4*>(row);
            // whole trips of 32 x uint4 while the row is complete (its pads reach a multiple of 128 keys)
            const bool whole = na == ncol;
            const int trips = rowlen >> 7;
            for (; bit >= stop; --bit) {
                const unsigned cand = key | (1u << bit);
                unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                if (whole) {
#pragma unroll 4
                    for (int j = 0; j < trips; ++j) {
                        const uint4 k = row4[lane + 32 * j];
                        c0 += k.x < cand;
                        c1 += k.y < cand;
                        c2 += k.z < cand;
                        c3 += k.w < cand;
                    }
                } else {
                    for (int i = lane; i < n4; i += 32) {
                        const uint4 k = row4[i];
                        c0 += k.x < cand;
                        c1 += k.y < cand;
                        c2 += k.z < cand;
                        c3 += k.w < cand;
                    }
                }
                const unsigned cnt = base + __reduce_add_sync(0xffffffffu, (c0 + c1) + (c2 + c3));
                if (cnt <= klo) { key = cand; below = cnt; }
            }
            // keep the keys that equal `key` in the bits above `bit`; the ones above that range only matter as
            // "the next larger key", the ones below it are counted in `below`
            const int sh = bit + 1;  // <= 28 (a level decides at least four bits), 0 after the last one
            const unsigned want = key >> sh;
            unsigned nc = 0;
            int i0 = 0;
            // four groups of 32 keys per trip: their loads are in flight together; all of them are in registers
            // before the first survivor is written (writes stay below i0 + 128)
            for (; i0 + 128 <= na; i0 += 128) {
                unsigned k[4], m[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) k[g] = row[i0 + 32 * g + lane];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const unsigned top = k[g] >> sh;
                    if (top > want) nxt = min(nxt, k[g]);
                    m[g] = __ballot_sync(0xffffffffu, top == want);
                }
                if ((m[0] | m[1] | m[2] | m[3]) == 0u) continue;  // the common case once the prefix is long
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if ((m[g] >> lane) & 1u) row[nc + __popc(m[g] & lt)] = k[g];
                    nc += __popc(m[g]);
                }
            }
            for (; i0 < na; i0 += 32) {
                const int i = i0 + lane;
                const unsigned k = (i < na) ? row[i] : 0u;
                const unsigned top = k >> sh;
                const bool hit = i < na && top == want;
                if (i < na && top > want) nxt = min(nxt, k);
                const unsigned m = __ballot_sync(0xffffffffu, hit);  // (every lane has read its key before anyone overwrites the row)
                if (hit) row[nc + __popc(m & lt)] = k;
                nc += __popc(m);
            }
            __syncwarp();
            na = (int)nc;
            base = below;
        }
        if (!done) {
            // every bit decided by counting: the survivors all equal key
            cnt_le = below + (unsigned)na;
            nxt = __reduce_min_sync(0xffffffffu, nxt);
        }
    }
    float m = key2f(key);
    if ((ncol & 1) == 0) {
        // upper middle: the same value if it is repeated, else the smallest key above it
        const float hi = (cnt_le >= klo + 2u) ? m : key2f(nxt);
        m = (m + hi) * 0.5f;  // numpy: mean of the two middle values in float32
    }
    if (lane == 0) {
        const size_t o = (size_t)sub * nfft + bin;
        if (med_lin) med_lin[o] = m;
        if (med_db) med_db[o] = power_to_db(m, eps);
    }
}

// ---- viewer-side reductions on the finished image ----------------------
// Minimum and maximum over the time axis, the two spectra proc_data's docstring promises next to the
// median (drfProc.py:430-433).  img [nsub][ncol][nfft]; a CTA owns 32 adjacent bins x 8 column slices,
// rows are read coalesced, slices meet in shared memory.  NaNs propagate like np.min / np.max.
__global__ void __launch_bounds__(256) minmax_time_kernel(const float* __restrict__ img, int nsub, int ncol, int nfft,
                                                          float eps, float* min_lin, float* max_lin, float* min_db,
                                                          float* max_db) {
    __shared__ float smin[8][32], smax[8][32];
    const int b = threadIdx.x & 31, s = threadIdx.x >> 5;
    const int blocks_per_sub = (nfft + 31) / 32;
    const int sub = blockIdx.x / blocks_per_sub;
    const int bin = (blockIdx.x - sub * blocks_per_sub) * 32 + b;
    const bool live = bin < nfft;
    const float* p = img + (size_t)sub * ncol * nfft + (live ? bin : nfft - 1);
    float lo = INFINITY, hi = -INFINITY;
    bool nan = false;
    for (int c = s; c < ncol; c += 8) {
        const float v = __ldg(p + (size_t)c * nfft);
        nan |= (v != v);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    if (nan) lo = hi = __int_as_float(0x7fc00000);
    smin[s][b] = lo;
    smax[s][b] = hi;
    __syncthreads();
    if (s == 0 && live) {
#pragma unroll
        for (int q = 1; q < 8; ++q) {
            const float a = smin[q][b], c2 = smax[q][b];
            if (a != a) lo = hi = a;
            else if (lo == lo) { lo = fminf(lo, a); hi = fmaxf(hi, c2); }
        }
        const size_t o = (size_t)sub * nfft + bin;
        if (min_lin) min_lin[o] = lo;
        if (max_lin) max_lin[o] = hi;
        if (min_db) min_db[o] = power_to_db(lo, eps);
        if (max_db) max_db[o] = power_to_db(hi, eps);
    }
}

// The bins the viewer actually draws: out[s][c][j] = clamp(img[s][c][idx[j]], lo, hi).  idx is the
// viewer's plotindices list (frequency-range selection and decimation to at most 2^15 points,
// drfview.py:1005-1023); the clamp is the colour-range clip of the PNG export (drfview.py:1515-1518;
// lo > hi disables it).  Shrinks the device-to-host copy to what is plotted.
__global__ void __launch_bounds__(256) gather_bins_kernel(const float* __restrict__ img, size_t rows, int nfft,
                                                          const int* __restrict__ idx, int count, float lo, float hi,
                                                          float* __restrict__ out) {
    const bool clamp = lo <= hi;
    const size_t total = rows * (size_t)count;
    for (size_t gi = blockIdx.x * (size_t)blockDim.x + threadIdx.x; gi < total; gi += (size_t)gridDim.x * blockDim.x) {
        const size_t r = gi / count;
        const int j = (int)(gi - r * count);
        float v = __ldg(img + r * nfft + __ldg(idx + j));
        if (clamp) {  // spectra[spectra < lo] = lo; spectra[spectra > hi] = hi  (NaN stays NaN)
            if (v < lo) v = lo;
            if (v > hi) v = hi;
        }
        out[gi] = v;
    }
}

// ---- large nfft (N = R0 * 4096, R0 = 2..16): two phases through an L2-resident scratch -----------
// A frame of N >= 16384 points does not fit the shared memory of one SM together with its
// pipeline (65536 points are 512 KB).  The first radix-R0 pass therefore runs as its own streaming
// kernel: thread n' loads x[n0*N2 + n'] (n0 < R0, N2 = N/R0 = 4096; coalesced across n'), applies
// the window, does the R0-point DFT in registers, multiplies output k0 by W_N^{n'*k0} and stores
// y[frame][k0][n'] to a scratch buffer sized to stay in the 126 MB L2.  The R0 sub-sequences
// y[frame][k0][:] are independent 4096-point transforms whose outputs are the bins k0 + R0*k', so
// phase B is the tuned 4096-point fused kernel run on the scratch with k0 in the role of the
// sub-channel, and phase C interleaves the R0 partial spectra into the fftshifted column.
struct SplitArgs {
    const void* iq;
    int iq_type;
    long long sample_stride, sub_stride, hop_elems;
    const long long* col_off;
    int ncol;         // columns per sub-channel (global)
    int cs_lo;        // first (sub*ncol + col) of this chunk
    int ncs_chunk;    // column-subchannel pairs in this chunk
    int k_lo, nfr_chunk;  // frame range [k_lo, k_lo + nfr_chunk) of every column in the chunk
    int col_frames;       // scratch frames reserved per column (>= nfr_chunk)
    const float* win;     // [N] w/sum(w)
    const float2* twa;    // [(R0-1)][N2] W_N^{n'*k0}, k0 = 1..R0-1
    float2* scratch;      // [ncs_chunk][col_frames][R0][N2]
};

PSG_DEV void stg_keep(float2* p, float2 v) {
    asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

template <int R0, int FPB>
__global__ void __launch_bounds__(256) sti_split_pass_kernel(const SplitArgs a) {
    constexpr int N2 = 4096;
    const int np = blockIdx.x * 256 + threadIdx.x;  // n'
    const int nframes = a.ncs_chunk * a.nfr_chunk;
    const int f0 = blockIdx.y * FPB;
    float w[R0];
    cf tw[R0 - 1];
#pragma unroll
    for (int n = 0; n < R0; ++n) w[n] = __ldg(a.win + n * N2 + np);
#pragma unroll
    for (int k = 1; k < R0; ++k) tw[k - 1] = __ldg(a.twa + (k - 1) * N2 + np);
#pragma unroll 1
    for (int f = f0; f < min(f0 + FPB, nframes); ++f) {
        const int csl = f / a.nfr_chunk, kk = f - csl * a.nfr_chunk;
        const int cs = a.cs_lo + csl;
        const int col = cs % a.ncol, sub = cs / a.ncol;
        const long long src = a.col_off[col] + (long long)sub * a.sub_stride + (long long)(a.k_lo + kk) * a.hop_elems;
        cf x[R0];
#pragma unroll
        for (int n = 0; n < R0; ++n) x[n] = ldg_iq_rt(a.iq_type, a.iq, src + (long long)(n * N2 + np) * a.sample_stride);
#pragma unroll
        for (int n = 0; n < R0; ++n) x[n] = cscale(x[n], w[n]);
        dftR<R0>(x);
        float2* dst = a.scratch + ((size_t)csl * a.col_frames + kk) * (R0 * N2) + np;
        stg_keep(dst, x[0]);
#pragma unroll
        for (int k = 1; k < R0; ++k) stg_keep(dst + (size_t)k * N2, cmul(x[k], tw[k - 1]));
    }
}

// Phase C: tmp[k0][c][(k' + N2/2) & (N2-1)] (the fftshifted sub-spectra phase B wrote, raw sums)
// -> column c, output index (k0 + R0*k' + N/2) mod N.  One CTA moves KT consecutive k' of one column
// through shared memory so that both sides are coalesced.  Columns whose frames span several
// chunks accumulate in `acc` (first: assign, later: add) and are scaled / converted on the last.
struct InterleaveArgs {
    const float* tmp;
    int r0, ncs_chunk, cs_lo;
    int first, last;
    float scale, eps;
    float* acc;      // [ncs_chunk][N] raw sums carried between frame blocks (null when single block)
    float* out_lin;  // [ncs_total][N] or null
    float* out_db;
};

__global__ void __launch_bounds__(256) sti_interleave_kernel(const InterleaveArgs a) {
    constexpr int N2 = 4096, KT = 128;
    __shared__ float tile[16][KT + 1];
    const int r0 = a.r0;
    const int N = r0 * N2;
    const int c = blockIdx.y;
    const int kp0 = blockIdx.x * KT;
    for (int e = threadIdx.x; e < r0 * KT; e += 256) {
        const int k0 = e / KT, j = e - k0 * KT;
        tile[k0][j] = a.tmp[((size_t)k0 * a.ncs_chunk + c) * N2 + (((kp0 + j) + N2 / 2) & (N2 - 1))];
    }
    __syncthreads();
    const size_t ocol = (size_t)(a.cs_lo + c) * N;
    for (int e = threadIdx.x; e < r0 * KT; e += 256) {
        const int j = e / r0, k0 = e - j * r0;
        const int k = k0 + r0 * (kp0 + j);
        const int idx = (k + N / 2) & (N - 1);
        float v = tile[k0][j];
        if (!a.first) v += a.acc[(size_t)c * N + idx];
        if (!a.last) {
            a.acc[(size_t)c * N + idx] = v;
        } else {
            const float p = v * a.scale;
            if (a.out_lin) a.out_lin[ocol + idx] = p;
            if (a.out_db) a.out_db[ocol + idx] = power_to_db(p, a.eps);
        }
    }
}

// ---- arbitrary (non power-of-two) nfft: Bluestein's chirp-z algorithm ----------------------------
// The viewer allows any integer FFT length 32..2^20 (drfview.py:474-479) and scipy's FFT takes them
// all.  With c[n] = exp(+j*pi*n^2/N):  X[k] = conj(c[k]) * sum_n (x[n] w[n] conj(c[n])) c[k-n], so
// |X[k]|^2 = |(a (*) c)[k]|^2 with a[n] = x[n] w[n] conj(c[n]) -- the unit-modulus post-chirp drops
// out of the power.  The circular convolution of length M = 2^m >= 2N-1 runs as
// IFFT_M(FFT_M(a) .* B), B = FFT_M(wrapped chirp)/M precomputed in float64 on the host.  Forward
// transform: radix-2 DIF (natural in, bit-reversed out); B is stored bit-reversed; inverse: radix-2
// DIT (bit-reversed in, natural out), so no reordering pass is needed.  Simple shared-memory
// radix-2 code: functional coverage of the reference's full nfft range, not a tuned path.
struct BluesteinArgs {
    const float2* aw;   // [N]  w[n]/sum(w) * conj(c[n])
    const float2* bbr;  // [M]  FFT_M(b)[bitrev(i)] / M
    const float2* twm;  // [M/2] exp(-2*pi*j*m/M)
    int n, logm;
};

__global__ void __launch_bounds__(256) sti_bluestein_kernel(const StiArgs a, const BluesteinArgs b, float2* gwork,
                                                            float* gacc) {
    const int N = b.n, M = 1 << b.logm;
    extern __shared__ __align__(16) float2 smem[];
    float2* buf = gwork ? gwork + (size_t)blockIdx.x * M : smem;
    float* accs = gacc ? gacc + (size_t)blockIdx.x * N : reinterpret_cast<float*>(smem + M);
    const int t = threadIdx.x, nt = blockDim.x;
    const int ncs = a.ncol * a.nsub;
    for (int item = blockIdx.x; item < ncs * a.nsplit; item += gridDim.x) {
        const int split = item % a.nsplit;
        const int cs = item / a.nsplit;
        const int col = cs % a.ncol, sub = cs / a.ncol;
        const int k0 = split * a.chunk;
        const int k1 = min(a.nfr, k0 + a.chunk);
        long long src = a.col_off[col] + (long long)sub * a.sub_stride + (long long)k0 * a.hop_elems;
        __syncthreads();
        for (int i = t; i < N; i += nt) accs[i] = 0.f;
        for (int k = k0; k < k1; ++k, src += a.hop_elems) {
            __syncthreads();
            for (int i = t; i < M; i += nt)
                buf[i] = (i < N) ? cmul(ldg_iq_rt(a.iq_type, a.iq, src + (long long)i * a.sample_stride), __ldg(b.aw + i))
                                 : make_float2(0.f, 0.f);
            for (int s = M >> 1; s >= 1; s >>= 1) {  // forward, DIF
                __syncthreads();
                const int step = (M >> 1) / s;
                for (int i = t; i < (M >> 1); i += nt) {
                    const int j = i & (s - 1);
                    const int base = ((i / s) * 2 * s) + j;
                    const cf u = buf[base], v = buf[base + s];
                    buf[base] = cadd(u, v);
                    const cf d = csub(u, v);
                    buf[base + s] = (s > 1) ? cmul(d, __ldg(b.twm + (size_t)j * step)) : d;
                }
            }
            __syncthreads();
            for (int i = t; i < M; i += nt) buf[i] = cmul(buf[i], __ldg(b.bbr + i));
            for (int s = 1; s <= (M >> 1); s <<= 1) {  // inverse, DIT, conjugate twiddles
                __syncthreads();
                const int step = (M >> 1) / s;
                for (int i = t; i < (M >> 1); i += nt) {
                    const int j = i & (s - 1);
                    const int base = ((i / s) * 2 * s) + j;
                    const cf u = buf[base];
                    cf v = buf[base + s];
                    if (s > 1) v = cmulc(v, __ldg(b.twm + (size_t)j * step));
                    buf[base] = cadd(u, v);
                    buf[base + s] = csub(u, v);
                }
            }
            __syncthreads();
            for (int i = t; i < N; i += nt) {
                const cf v = buf[i];
                accs[i] = fmaf(v.x, v.x, fmaf(v.y, v.y, accs[i]));
            }
        }
        __syncthreads();
        const int half = N / 2;  // np.fft.fftshift: out[(k + N//2) mod N] = in[k]
        for (int i = t; i < N; i += nt) {
            int idx = i + half;
            if (idx >= N) idx -= N;
            if (a.nsplit > 1) {
                a.partial[((size_t)cs * a.nsplit + split) * N + idx] = accs[i];
            } else {
                const float p = accs[i] * a.scale;
                if (a.out_lin) a.out_lin[(size_t)cs * N + idx] = p;
                if (a.out_db) a.out_db[(size_t)cs * N + idx] = power_to_db(p, a.eps);
            }
        }
    }
}
