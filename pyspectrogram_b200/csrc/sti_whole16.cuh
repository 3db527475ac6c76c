// nfft = 16384 with the whole frame in one SM's shared memory and THREE shared-memory round trips per
// sample instead of the four of sti_whole.cuh:
//
//   plan      N = 16 x 2 x 16 x 2 x 16, strides 1024, 512, 32, 16, 1 (index algebra of sti_kernels.cuh).  Both
//             radix-2 passes run in registers on the end of the radix-16 pass before them: the two inputs of a
//             radix-2 butterfly sit in lanes l and l ^ 16 of one warp, which swap half of their 16 outputs by
//             SHFL and finish the butterflies of the half they keep (pair_exchange_store, sti_kernels.cuh).
//   pass 0+1  radix 16 over n0 (elements n' + n0*1024), fed in four slabs.  Slab m holds n' = n'' + 512 hi for
//             n'' in [128 m, 128 m + 128), hi = 0, 1: 32 bulk copies (UBLKCP) of 1 KB into one 32 KB stage.  The
//             CTA's two halves of 256 threads alternate slabs (half h takes m = h, h + 2) and each owns one stage:
//             lane l of warp w8 has n'' = 128 m + 16 w8 + (l & 15), hi = l >> 4, so its radix-2 partner is lane
//             l ^ 16.  Windowed 16-point DFT, W_N^{n' k0} rebuilt from the four powers W_N^{n' 2^q} (table loads issued before
//             the wait for the slab), pair exchange with W_1024^{n''}, stores to
//             k0*1024 + k1*512 + n''.
//   pass 2+3  radix 16 at stride 32 inside each 512-point block plus the radix-2 pass at stride 16
//             (the pass of the 8192-point 16 x 16 x 2 x 16 kernel), two butterflies per thread, in place.
//   pass 4    radix 16 on consecutive positions, |X|^2 into 32 accumulators per thread.  A warp reads back
//             exactly the 512-point blocks it wrote in pass 2+3: __syncwarp.
// 7 shared-memory accesses per sample (bulk-copy write, stage read, 2 x (write + read), last read, 2 x half an
// access worth of SHFL) against 8, and two CTA barriers per frame against three.  512 threads, one CTA per SM.
// The stage of a half is refilled right after the CTA barrier that follows its reads, every warp issuing four
// of the 32 copies, so the first slab of frame f+1 streams in under passes 2-4 of frame f.
#pragma once
#include "sti_whole.cuh"

template <int IQT>
struct Whole16Cfg {
    static constexpr int N = 16384, T = 512, WS = 128, NSL = 4;
    static constexpr int IQB = IqBytes<IQT>::value;
    static constexpr int SEG = WS * IQB + 16;  // one staged run of 128 samples + alignment slack
    static constexpr int STAGE = 32 * SEG;     // 16 n0 x 2 halves
    static constexpr int NPAD = psg_pad(N) + 2;
    static constexpr int HDR = 128;
    static constexpr size_t smem_bytes = HDR + 2 * (size_t)STAGE + (size_t)NPAD * 8;
};

struct Whole16Args {
    StiArgs s;  // tw = full table W_N^m
};

// pair exchange after the first pass: as pair_exchange_store with the strides of the 1024 / 512 split
PSG_DEV void pair_exchange_store_1024(const cf* a, float2* q, const bool hi, const cf w1) {
    if (hi) q += 8 * pad_off(1024);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const cf send = hi ? a[i] : a[8 + i];
        const cf keep = hi ? a[8 + i] : a[i];
        cf recv;
        recv.x = __shfl_xor_sync(0xffffffffu, send.x, 16);
        recv.y = __shfl_xor_sync(0xffffffffu, send.y, 16);
        q[i * pad_off(1024)] = cadd(keep, recv);
        q[i * pad_off(1024) + pad_off(512)] = cmul(csub(keep, recv), w1);
    }
}

template <int IQT>
__global__ void __launch_bounds__(512, 1) sti_whole16_kernel(const Whole16Args wa) {
    using CF = Whole16Cfg<IQT>;
    constexpr int N = CF::N, T = CF::T, IQB = CF::IQB, SEG = CF::SEG;
    const StiArgs& a = wa.s;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);      // [2] stage of half h full
    unsigned char* stage = smem_raw + CF::HDR;
    float2* xch = reinterpret_cast<float2*>(smem_raw + CF::HDR + 2 * (size_t)CF::STAGE);

    const int t = threadIdx.x;
    const int h = t >> 8;                    // half of the CTA: slabs h and h + 2
    const int lane = t & 31;
    const bool hi = (lane & 16) != 0;        // n' = n'' + 512 hi
    const int nt = 16 * ((t & 255) >> 5) + (lane & 15);  // n'' - 128 m
    const int item = blockIdx.x;
    const int split = item % a.nsplit;
    const int cs = item / a.nsplit;
    const int col = cs % a.ncol, sub = cs / a.ncol;
    const int kf0 = split * a.chunk;
    const int nfr = min(a.nfr, kf0 + a.chunk) - kf0;
    const long long fbase = a.col_off[col] + (long long)sub * a.sub_stride + (long long)kf0 * a.hop_elems;
    const int nsteps = 2 * nfr;  // per half: step q = (frame q >> 1, slab h + 2 (q & 1))
    unsigned char* const myst = stage + (size_t)h * CF::STAGE;

    // Refill of this half's stage for step q, spread over the half's eight warps: lane 0 of warp w8 issues
    // copies 4 w8 .. 4 w8 + 3, warp 0 also posts the byte count.  (One thread issuing all 32 copies is a
    // ~450-instruction serial section in front of a CTA barrier: 14 % of all stall samples.)  Called right
    // after a CTA barrier that follows every warp's reads of the stage, so no "stage free" signal is needed;
    // a copy that completes before the count is posted only makes the pending byte count negative for a while.
    const int w8 = (t & 255) >> 5;
    auto issue = [&](int q) {  // lane 0 of every warp
        const int f = q >> 1, m = h + 2 * (q & 1);
        const uintptr_t src0 =
            reinterpret_cast<uintptr_t>(a.iq) + (uintptr_t)((fbase + (long long)f * a.hop_elems + m * CF::WS) * IQB);
        const uint32_t bytes = CF::WS * IQB + ((src0 & 15) ? 16 : 0);
        uint64_t* bar = bars + h;
        if (w8 == 0) mbar_expect_tx(bar, bytes * 32);
        const uintptr_t s16 = (src0 & ~(uintptr_t)15) + (uintptr_t)(4 * w8) * 512 * IQB;
        unsigned char* dst = myst + 4 * w8 * SEG;
#pragma unroll
        for (int c = 0; c < 4; ++c)  // copy 2 n0 + hi: element offset 512 (2 n0 + hi)
            bulk_g2s(dst + c * SEG, reinterpret_cast<const void*>(s16 + (uintptr_t)c * 512 * IQB), bytes, bar);
    };
    if (t == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // barriers and counters initialised
    if (lane == 0) issue(0);

    // (twiddles are re-read from L1/L2 right before the wait or barrier that precedes their use: as
    // loop-invariant registers they push the accumulators into local memory)
    const int ct = nt + 128 * h + (hi ? 512 : 0);  // n' of this thread's first slab; the second is 256 further
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;

    // window of the thread's 16 samples of a slab (L2 round trips: requested well before their use)
    float wq[16];
    auto load_window = [&](int j) {
        const float* wp = a.win + (hi ? 512 : 0) + 128 * (h + 2 * j) + nt;
#pragma unroll
        for (int n0 = 0; n0 < 16; ++n0) wq[n0] = __ldg(wp + n0 * 1024);
    };
    load_window(0);
    float2* const q0 = xch + psg_pad(nt + 128 * h);

    for (int f = 0; f < nfr; ++f) {
        const int skew =
            (int)(((reinterpret_cast<uintptr_t>(a.iq) + (uintptr_t)((fbase + (long long)f * a.hop_elems) * IQB)) & 15) / IQB);
        // ---- pass 0+1: this half's two slabs ----
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
            const int q = 2 * f + j;
            cf x[16];
            cf pw[4];  // W_N^{n' 2^q}
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) pw[qq] = __ldg(a.tw + ((ct + 256 * j) << qq));
            cf w1 = __ldg(a.tw + 16 * (nt + 128 * h + 256 * j));  // W_1024^{n''}
            if (hi) w1 = make_float2(-w1.x, -w1.y);
            mbar_wait(bars + h, q & 1);
#pragma unroll
            for (int n0 = 0; n0 < 16; ++n0) x[n0] = lds_iq<IQT>(myst + (2 * n0 + (hi ? 1 : 0)) * SEG, skew + nt);
            dftRw<16>(x, wq);
            if (j == 0) load_window(1);
            twiddle_dfs<16>(x, pw);
            if (j == 0) {
                // the last pass of the previous frame is done with the exchange buffer, and every warp's
                // butterflies have consumed its loads from the stage: refill it with the second slab
                __syncthreads();
                if (lane == 0) issue(q + 1);
            }
            pair_exchange_store_1024(x, q0 + pad_off(256 * j), hi, w1);
        }
        cf w2 = __ldg(a.tw + 512 * (lane & 15));  // W_32^{lane & 15}
        if (hi) w2 = make_float2(-w2.x, -w2.y);
        cf pwA[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) pwA[q] = __ldg(a.tw + ((32 * lane) << q));  // W_512^{lane 2^q}
        __syncthreads();
        if (lane == 0 && 2 * f + 2 < nsteps) issue(2 * f + 2);  // next frame's first slab streams in under passes 2-4
        // ---- pass 2+3: radix 16 at stride 32 + radix 2 at stride 16, inside each 512-point block ----
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
            const int b = t + i * T;
            float2* p = xch + psg_pad((b >> 5) * 512 + lane);
            cf v[16];
#pragma unroll
            for (int n = 0; n < 16; ++n) v[n] = p[pad_off(n * 32)];
            dftR<16>(v);
            twiddle_dfs<16>(v, pwA);
            pair_exchange_store(v, xch + psg_pad((b >> 5) * 512 + (lane & 15)), hi, w2);
        }
        __syncwarp();  // a warp reads back the two 512-point blocks it wrote
        if (f + 1 < nfr) load_window(0);  // next frame's first slab, in flight under the last pass
        // ---- pass 4: radix 16 on consecutive positions, |X|^2 into the accumulators ----
        smem_pass<32, T, 16, 1, true>(xch, nullptr, t, acc);
    }

    // ---- epilogue: digit-reversed register sums -> fftshifted, coalesced 128-bit stores ----
    __syncthreads();
    float* sout = reinterpret_cast<float*>(xch);  // N floats
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        // last-pass butterfly b = k0*64 + k1*32 + k2*2 + k3 holds frequencies k0 + 16 k1 + 32 k2 + 512 k3 + 1024 j
        const int b = t + i * T;
        const int klow = (b >> 6) + 16 * ((b >> 5) & 1) + 32 * ((b >> 1) & 15) + 512 * (b & 1);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int freq = klow + 1024 * jj;
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
