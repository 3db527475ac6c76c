// Large nfft (N = R0 * 4096, R0 = 2..16) as ONE kernel: a thread-block cluster of R0 CTAs per frame.
//
// A frame of N >= 16384 complex64 points (128..512 KB) does not fit one SM's shared memory together
// with its pipeline, so the transform is split four-step style, N = R0 x 4096 with n = n0*4096 + n',
// k = k0 + R0*k':
//   X[k0 + R0 k'] = sum_{n'} W_4096^{n' k'} * ( W_N^{n' k0} * sum_{n0} w[n] x[n] W_R0^{n0 k0} )
// CTA c of the cluster
//   pre-pass   owns the slab n' in [c*4096/R0, (c+1)*4096/R0): its R0 segments of the frame arrive by
//              TMA bulk copies (two-stage ring, one frame ahead), the thread does the windowed R0-point
//              DFTs over n0 in registers, applies W_N^{n' k0} and stores output k0 to row k0 of the
//              cluster's exchange slot in global memory -- 1.5 MB per cluster, rewritten every third
//              frame, so it lives in the 126 MB L2 and never reaches HBM;
//   row pass   owns row k0 = c: after the cluster barrier it reads the 4096 points of its row back
//              (coalesced, L2 hits) straight into registers and runs the tuned 16x16x16 transform of
//              the 4096-point kernels; |X|^2 of bins c + R0*k' accumulate in registers over the frames.
// The exchange goes through L2 and not through distributed shared memory because DSMEM moves
// ~20 B/clk/SM (B300_MICROARCH.md), less than one SM's share of HBM; only the synchronisation uses
// the cluster: one hardware barrier per frame (barrier.cluster arrive.release / wait.acquire), split
// so that the pre-pass of frame f+1 runs between the arrive and the wait of frame f.  The exchange
// slot is triple-buffered, which is what makes one barrier per frame sufficient (see the loop).
// Clusters are persistent: cluster q walks work items q, q + nclusters, ... (item = one column's
// chunk of frames) as one continuous pipeline; finished items are written as raw sums
// tmp[item][k0][k'] and sti_cluster_finalize_kernel interleaves, fftshifts, scales and converts.
// HBM sees every sample once: 8 B/sample in, 4 B/bin per item out.
#pragma once
#include "sti_kernels.cuh"

struct ClusterArgs {
    const void* iq;            // contiguous samples (sample_stride == 1), 16-byte aligned base
    long long sub_stride;      // elements
    long long hop_elems;
    const long long* col_off;  // [ncol]
    int ncol, ncs;             // columns per sub-channel, ncol * nsub
    int nfr, chunk, nsplit;    // frames per column, frames per item, items per column
    int nclusters;
    const float* win;          // [N] w/sum(w)
    const float2* twa;         // [R0-1][4096]  W_N^{n'*k0}, k0 = 1..R0-1
    const float2* twp;         // pass tables of the 4096-point 16x16x16 plan (power layout, TWP = 2)
    float2* scratch;           // [nclusters][3][R0][4096]
    float* tmp;                // [ncs*nsplit][R0][4096] raw sums, natural k'
};

// exchange-slot read: written by other CTAs of the cluster during this kernel -> not the .nc path;
// .cg keeps it out of L1 (the acquire of the cluster barrier orders it after the writers' release)
PSG_DEV float2 ld_xchg(const float2* p) {
    float2 v;
    asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
// W^k, k = 1..R-1, from W^1, W^2, W^4, W^8 (one complex multiply per non-power)
template <int R>
PSG_DEV void rebuild_tw(const cf* pw, cf* tw) {
#pragma unroll
    for (int k = 1; k < R; ++k) {
        const int q = (k >= 8) ? 3 : (k >= 4) ? 2 : (k >= 2) ? 1 : 0;
        const int hb = 1 << q;
        tw[k - 1] = (k == hb) ? pw[q] : cmul(tw[k - hb - 1], pw[q]);
    }
}

// x[k] *= W^k, k = 1..R-1, with W^k built on the fly from the powers pw[q] = W^(2^q) in depth-first
// order: k = 2^q0 + 2^q1 + .. (q0 < q1 < ..) is ((pw[q0]*pw[q1])*pw[q2])*pw[q3] -- the association of
// rebuild_tw, so both give identical values -- but every product is consumed at once, so at most
// three of them are live instead of fifteen.
template <int R, int K, int QMIN>
PSG_DEV void tw_visit(cf* x, const cf* pw, const cf tk) {
    x[K] = cmul(x[K], tk);
    if constexpr (QMIN <= 0 && K + 1 < R) tw_visit<R, K + 1, 1>(x, pw, cmul(tk, pw[0]));
    if constexpr (QMIN <= 1 && K + 2 < R) tw_visit<R, K + 2, 2>(x, pw, cmul(tk, pw[1]));
    if constexpr (QMIN <= 2 && K + 4 < R) tw_visit<R, K + 4, 3>(x, pw, cmul(tk, pw[2]));
    if constexpr (QMIN <= 3 && K + 8 < R) tw_visit<R, K + 8, 4>(x, pw, cmul(tk, pw[3]));
}
template <int R>
PSG_DEV void twiddle_dfs(cf* x, const cf* pw) {
    if constexpr (R > 1) tw_visit<R, 1, 1>(x, pw, pw[0]);
    if constexpr (R > 2) tw_visit<R, 2, 2>(x, pw, pw[1]);
    if constexpr (R > 4) tw_visit<R, 4, 3>(x, pw, pw[2]);
    if constexpr (R > 8) tw_visit<R, 8, 4>(x, pw, pw[3]);
}

struct ClusterStep {  // one frame of the cluster's pipeline, published by the producer thread
    int valid, last, item, skew;
};

template <int R0, int IQT>
struct ClusterCfg {
    static constexpr int N2 = 4096, E = 16, T = 256;
    static constexpr int N = R0 * N2;
    static constexpr int SL = N2 / R0;   // slab width (n' per CTA)
    static constexpr int NBA = E / R0;   // pre-pass butterflies per thread
    static constexpr int IQB = IqBytes<IQT>::value;
    static constexpr int SEG = SL * IQB + 16;  // staged segment + alignment slack
    static constexpr int STAGE = R0 * SEG;
    static constexpr int NPAD = psg_pad(N2) + 2;
    static constexpr int HDR = 128;  // 2 mbarriers, producer cursor, step ring
    static constexpr int TW0S = 4 * 256 * 8;  // W_4096^{t*2^q}, q < 4 (TW0Q: row pass-0 twiddle powers)
    static constexpr size_t smem_bytes = HDR + TW0S + 2 * (size_t)STAGE + (size_t)NPAD * 8;
};

template <int R0, int IQT, int ROWTMA>
__global__ void __launch_bounds__(256, 2) sti_cluster_kernel(const ClusterArgs a) {
    using CF = ClusterCfg<R0, IQT>;
    constexpr int N2 = CF::N2, E = CF::E, T = CF::T, SL = CF::SL, NBA = CF::NBA, IQB = CF::IQB, SEG = CF::SEG;
    constexpr int NPWA = psg_npow(R0);
    using PL = Plan<N2, 16, 16, 16, 1, 2>;
    static_assert(PL::ROW1 && PL::S1 == 16 && PL::S0 == 256, "4096 = 16*16*16 plan");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);                 // [3]: slab stages 0/1, row
    long long* cur_fb = reinterpret_cast<long long*>(smem_raw + 24);        // producer cursor (thread 0 only)
    int* cur = reinterpret_cast<int*>(smem_raw + 32);                       // item, k, k1, step index
    ClusterStep* ring = reinterpret_cast<ClusterStep*>(smem_raw + 64);      // [4]
    float2* tw0s = reinterpret_cast<float2*>(smem_raw + CF::HDR);           // [4][256]
    unsigned char* stage = smem_raw + CF::HDR + CF::TW0S;
    float2* xch = reinterpret_cast<float2*>(smem_raw + CF::HDR + CF::TW0S + 2 * (size_t)CF::STAGE);

    const int t = threadIdx.x;
    const int c = (int)cluster_ctarank();
    const int cid = blockIdx.x / R0;
    const int nitems = a.ncs * a.nsplit;

    // ---- producer (thread 0): walks the cluster's frames, publishes the step, issues its TMA ----
    auto item_setup = [&](int item) {
        const int cs = item / a.nsplit, split = item - cs * a.nsplit;
        const int col = cs % a.ncol, sub = cs / a.ncol;
        const int k0 = split * a.chunk;
        cur[0] = item;
        cur[1] = k0;
        cur[2] = min(a.nfr, k0 + a.chunk);
        *cur_fb = a.col_off[col] + (long long)sub * a.sub_stride + (long long)k0 * a.hop_elems;
    };
    auto produce = [&]() {  // thread 0 only
        const int s = cur[3]++;
        ClusterStep st;
        st.item = cur[0];
        st.valid = st.item < nitems;
        st.last = 0;
        st.skew = 0;
        if (st.valid) {
            const long long fb = *cur_fb;
            const uintptr_t src0 = reinterpret_cast<uintptr_t>(a.iq) + (uintptr_t)((fb + (long long)c * SL) * IQB);
            const uint32_t mis = (uint32_t)(src0 & 15);
            st.skew = (int)(mis / IQB);
            const uint32_t bytes = SL * IQB + (mis ? 16 : 0);
            uint64_t* bar = bars + (s & 1);
            mbar_expect_tx(bar, bytes * R0);
            unsigned char* dst = stage + (size_t)(s & 1) * CF::STAGE;
#pragma unroll 1
            for (int n0 = 0; n0 < R0; ++n0)
                bulk_g2s(dst + n0 * SEG, reinterpret_cast<const void*>((src0 & ~(uintptr_t)15) + (uintptr_t)n0 * N2 * IQB), bytes,
                         bar);
            const int k = ++cur[1];
            st.last = k >= cur[2];
            if (st.last) {
                const int nx = st.item + a.nclusters;
                if (nx < nitems) item_setup(nx);
                else cur[0] = nitems;
            } else {
                *cur_fb = fb + a.hop_elems;
            }
        }
        ring[s & 3] = st;
    };
    if (t == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        mbar_init(bars + 2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        cur[3] = 0;
        if (cid < nitems) item_setup(cid);
        else cur[0] = nitems;
        produce();
        produce();
    }

    // ---- loop-invariant tables in registers ----
    float w[E];         // window of this thread's pre-pass samples
    cf pwa[NBA * NPWA]; // W_N^{n' * 2^q} of this thread's pre-pass columns
#pragma unroll
    for (int i = 0; i < NBA; ++i) {
        const int np = c * SL + t + i * T;
#pragma unroll
        for (int n0 = 0; n0 < R0; ++n0) w[i * R0 + n0] = __ldg(a.win + n0 * N2 + np);
#pragma unroll
        for (int q = 0; q < NPWA; ++q) pwa[i * NPWA + q] = __ldg(a.twa + ((1 << q) - 1) * N2 + np);
    }
    // row pass 0 twiddle powers W_4096^{t*2^q} live in shared memory: with them in registers the
    // kernel spills (ptxas: 152 bytes), and four LDS.64 per frame cost less than the spill traffic
#pragma unroll
    for (int q = 0; q < 4; ++q) tw0s[q * 256 + t] = __ldg(a.twp + PL::TW0 + ((1 << q) - 1) * PL::S0 + t);
    cf wb1[4];
    {
        const int npr = t & (PL::S1 - 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) wb1[q] = __ldg(a.twp + PL::TW1 + npr * 6 + q);
    }
    float acc[E];
#pragma unroll
    for (int i = 0; i < E; ++i) acc[i] = 0.f;

    float2* const slot = a.scratch + (size_t)cid * 3 * CF::N;

    // pre-pass of step s: staged slab -> windowed R0-point DFTs -> twiddle -> exchange slot s % 3
    auto prepass = [&](int s, int skew) {
        mbar_wait(bars + (s & 1), (s >> 1) & 1);
        const unsigned char* sb = stage + (size_t)(s & 1) * CF::STAGE;
        float2* dst = slot + (size_t)(s % 3) * CF::N + c * SL + t;
#pragma unroll
        for (int i = 0; i < NBA; ++i) {
            cf x[R0];
#pragma unroll
            for (int n0 = 0; n0 < R0; ++n0) x[n0] = lds_iq<IQT>(sb + n0 * SEG, skew + t + i * T);
            dftRw<R0>(x, &w[i * R0]);
            twiddle_dfs<R0>(x, &pwa[i * NPWA]);
#pragma unroll
            for (int k0 = 0; k0 < R0; ++k0) stg_keep(dst + (size_t)k0 * N2 + i * T, x[k0]);
        }
        // the rows are read back by bulk copies (async proxy): order these generic-proxy stores before them
        if constexpr (ROWTMA) asm volatile("fence.proxy.async;" ::: "memory");
    };
    // raw sums of a finished item -> tmp[item][c][k'] (natural order), through shared memory
    auto epilogue = [&](int item) {
        __syncthreads();
        float* sout = reinterpret_cast<float*>(xch);
        const int klow = PL::low_freq(t);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int freq = klow + (N2 / 16) * jj;
            sout[freq ^ (((freq >> 5) & 7) << 2)] = acc[jj];
        }
        __syncthreads();
        const float4* sout4 = reinterpret_cast<const float4*>(sout);
        float4* dst = reinterpret_cast<float4*>(a.tmp + ((size_t)item * R0 + c) * N2);
#pragma unroll
        for (int q = t; q < N2 / 4; q += T) dst[q] = sout4[q ^ ((q >> 3) & 7)];
    };

    __syncthreads();  // barriers initialised, steps 0 and 1 published
    if (!ring[0].valid) return;  // cluster-uniform: no work for this cluster
    prepass(0, ring[0].skew);
    cluster_arrive();
    __syncthreads();  // stage 0 has been read by everyone
    if (t == 0) produce();  // step 2 -> stage 0

    // Iteration r = row transform of frame r, with the pre-pass of frame r+1 inside it.
    //   arrive #r+1 is issued after this CTA's pre-pass stores of frame r+1 AND its row read of frame r;
    //   wait #r (matching arrive #r of every CTA) therefore guarantees that all rows of frame r are
    //   written and that nobody still reads the slot the next pre-pass overwrites.
    // ROWTMA = 1: the row comes back as one 32 KB bulk copy into the (idle) exchange buffer, issued
    //   right after wait #r so that its L2 latency hides behind the pre-pass of frame r+1.
    // ROWTMA = 0: the pre-pass runs first (a full iteration of barrier slack, which is why the slot is
    //   triple-buffered) and the row is loaded straight into registers after the wait (latency exposed).
    for (int r = 0;; ++r) {
        const ClusterStep cur_step = ring[r & 3];
        if (!cur_step.valid) break;
        const ClusterStep nxt = ring[(r + 1) & 3];
        cf x[E];
        if constexpr (ROWTMA) {
            __syncthreads();  // the exchange buffer is idle: last pass / epilogue of frame r-1 are done with it
            cluster_wait();
            if (t == 0) {
                asm volatile("fence.proxy.async;" ::: "memory");
                mbar_expect_tx(bars + 2, N2 * 8);
                bulk_g2s(xch, slot + (size_t)(r % 3) * CF::N + (size_t)c * N2, N2 * 8, bars + 2);
            }
            if (nxt.valid) prepass(r + 1, nxt.skew);
            mbar_wait(bars + 2, r & 1);
#pragma unroll
            for (int n = 0; n < 16; ++n) x[n] = xch[t + n * PL::S0];
        } else {
            if (nxt.valid) prepass(r + 1, nxt.skew);
            cluster_wait();
            const float2* row = slot + (size_t)(r % 3) * CF::N + (size_t)c * N2 + t;
#pragma unroll
            for (int n = 0; n < 16; ++n) x[n] = ld_xchg(row + n * PL::S0);
        }
        dftR<16>(x);
        cluster_arrive();
        {
            cf pw[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) pw[q] = tw0s[q * 256 + t];
            twiddle_dfs<16>(x, pw);
        }
        __syncthreads();  // xch: frame r-1 (or the staged row) has been read; stage (r+1)&1 has been read
        if (t == 0) produce();  // step r+3 -> stage (r+1)&1
        {
            float2* p0 = xch + psg_pad(t);
#pragma unroll
            for (int k = 0; k < 16; ++k) p0[pad_off(k * PL::S0)] = x[k];
        }
        {
            cf tw1[15];
            rebuild_tw<16>(wb1, tw1);
            __syncthreads();
            smem_pass<E, T, 16, PL::S1, false>(xch, tw1, t, acc);
        }
        __syncwarp();  // 16x16 blocks of passes 1 and 2 stay inside aligned groups of 16 threads
        smem_pass<E, T, 16, 1, true>(xch, nullptr, t, acc);
        if (cur_step.last) {
            epilogue(cur_step.item);
#pragma unroll
            for (int i = 0; i < E; ++i) acc[i] = 0.f;
        }
    }
    cluster_wait();  // consume the last arrive
}

// ---- the same decomposition with the exchange in distributed shared memory ------------------------
// Every pre-pass output goes straight from registers into the row owner's exchange buffer with
// st.async (8 bytes, completion counted on the owner's mbarrier), already in the padded layout the
// 4096-point passes use, so the row transform runs in place where the data lands: no scratch in L2,
// no memory fences (the L2 variant spends ~30 % of its stall samples in MEMBAR.GPU behind the
// cluster-scope release), two CTA barriers per frame instead of four.
//   full[b]  (count 1 + 32 KB of transactions)  frame r has landed in this CTA's buffer r & 1
//   free[b]  (count R0)                          every CTA of the cluster is done with buffer b of
//                                                frame r, so frame r+2 may be sent
// Pre-pass of frame r+1 runs before the row transform of frame r: the data of a frame has a whole row
// phase to cross the cluster, only the small "free" signal is waited for with no slack.
PSG_DEV void mbar_arrive_cluster(uint32_t rbar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}

struct DsmemStep {  // ClusterStep + the address its slab is fetched from
    int valid, last, item, skew;
    unsigned long long src;
    unsigned long long pad_;
};

template <int R0, int IQT>
struct DsmemCfg {
    using CC = ClusterCfg<R0, IQT>;
    static constexpr int HDR = 192;  // 5 mbarriers, producer cursor, step ring
    static constexpr size_t XB = (size_t)CC::NPAD * 8;
    static constexpr size_t smem_bytes = HDR + (size_t)CC::STAGE + 2 * XB;
};

template <int R0, int IQT>
__global__ void __launch_bounds__(256, 2) sti_dsmem_kernel(const ClusterArgs a) {
    using CF = ClusterCfg<R0, IQT>;
    using DC = DsmemCfg<R0, IQT>;
    constexpr int N2 = CF::N2, E = CF::E, T = CF::T, SL = CF::SL, NBA = CF::NBA, IQB = CF::IQB, SEG = CF::SEG;
    constexpr int NPWA = psg_npow(R0);
    using PL = Plan<N2, 16, 16, 16, 1, 2>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar_slab = reinterpret_cast<uint64_t*>(smem_raw);            // [1]
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw) + 1;        // [2]
    uint64_t* bar_free = reinterpret_cast<uint64_t*>(smem_raw) + 3;        // [2]
    long long* cur_fb = reinterpret_cast<long long*>(smem_raw + 40);       // producer cursor (thread 0 only)
    int* cur = reinterpret_cast<int*>(smem_raw + 48);                      // item, k, k1, step index
    DsmemStep* ring = reinterpret_cast<DsmemStep*>(smem_raw + 64);       // [4]
    unsigned char* stage = smem_raw + DC::HDR;
    float2* xb0 = reinterpret_cast<float2*>(smem_raw + DC::HDR + CF::STAGE);

    const int t = threadIdx.x;
    const int c = (int)cluster_ctarank();
    const int cid = blockIdx.x / R0;
    const int nitems = a.ncs * a.nsplit;

    auto item_setup = [&](int item) {
        const int cs = item / a.nsplit, split = item - cs * a.nsplit;
        const int col = cs % a.ncol, sub = cs / a.ncol;
        const int k0 = split * a.chunk;
        cur[0] = item;
        cur[1] = k0;
        cur[2] = min(a.nfr, k0 + a.chunk);
        *cur_fb = a.col_off[col] + (long long)sub * a.sub_stride + (long long)k0 * a.hop_elems;
    };
    // thread 0 only.  publish(): walk the cluster's frames and describe the next step in the ring;
    // fetch(s): issue the bulk copies of step s into the (single) slab stage once it has been read.
    // A step is published three iterations before its row transform, i.e. two CTA barriers before
    // anybody reads it, and fetched one iteration before its pre-pass.
    auto publish = [&]() {
        const int s = cur[3]++;
        DsmemStep st;
        st.item = cur[0];
        st.valid = st.item < nitems;
        st.last = 0;
        st.skew = 0;
        st.src = 0;
        st.pad_ = 0;
        if (st.valid) {
            const long long fb = *cur_fb;
            const uintptr_t src0 = reinterpret_cast<uintptr_t>(a.iq) + (uintptr_t)((fb + (long long)c * SL) * IQB);
            st.skew = (int)((src0 & 15) / IQB);
            st.src = (unsigned long long)src0;
            const int k = ++cur[1];
            st.last = k >= cur[2];
            if (st.last) {
                const int nx = st.item + a.nclusters;
                if (nx < nitems) item_setup(nx);
                else cur[0] = nitems;
            } else {
                *cur_fb = fb + a.hop_elems;
            }
        }
        ring[s & 3] = st;
    };
    auto fetch = [&](int s) {
        const DsmemStep st = ring[s & 3];
        if (!st.valid) return;
        const uint32_t bytes = SL * IQB + ((st.src & 15) ? 16 : 0);
        mbar_expect_tx(bar_slab, bytes * R0);
#pragma unroll 1
        for (int n0 = 0; n0 < R0; ++n0)
            bulk_g2s(stage + n0 * SEG, reinterpret_cast<const void*>((st.src & ~15ull) + (unsigned long long)n0 * N2 * IQB), bytes,
                     bar_slab);
    };
    if (t == 0) {
        mbar_init(bar_slab, 1);
        mbar_init(bar_full + 0, 1);
        mbar_init(bar_full + 1, 1);
        mbar_init(bar_free + 0, R0);
        mbar_init(bar_free + 1, R0);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_full + 0, N2 * 8);  // frames 0 and 1
        mbar_expect_tx(bar_full + 1, N2 * 8);
        cur[3] = 0;
        if (cid < nitems) item_setup(cid);
        else cur[0] = nitems;
        publish();
        publish();
        publish();
        fetch(0);
    }
    // nobody may touch a peer's barriers or buffers before they are initialised
    cluster_arrive();
    cluster_wait();

    float w[E];
    cf pwa[NBA * NPWA];
#pragma unroll
    for (int i = 0; i < NBA; ++i) {
        const int np = c * SL + t + i * T;
#pragma unroll
        for (int n0 = 0; n0 < R0; ++n0) w[i * R0 + n0] = __ldg(a.win + n0 * N2 + np);
#pragma unroll
        for (int q = 0; q < NPWA; ++q) pwa[i * NPWA + q] = __ldg(a.twa + ((1 << q) - 1) * N2 + np);
    }
    cf wb1[4];
    {
        const int npr = t & (PL::S1 - 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) wb1[q] = __ldg(a.twp + PL::TW1 + npr * 6 + q);
    }
    float acc[E];
#pragma unroll
    for (int i = 0; i < E; ++i) acc[i] = 0.f;

    // pre-pass of step s: staged slab -> windowed R0-point DFTs -> twiddle -> row owners' buffer s & 1
    auto prepass = [&](int s, int skew) {
        mbar_wait_bounded(bar_slab, s & 1);
        const uint32_t dst = smem_u32(xb0) + (uint32_t)((s & 1) * DC::XB);
        const uint32_t fullb = smem_u32(bar_full + (s & 1));
#pragma unroll
        for (int i = 0; i < NBA; ++i) {
            cf x[R0];
#pragma unroll
            for (int n0 = 0; n0 < R0; ++n0) x[n0] = lds_iq<IQT>(stage + n0 * SEG, skew + t + i * T);
            dftRw<R0>(x, &w[i * R0]);
            twiddle_dfs<R0>(x, &pwa[i * NPWA]);
            // buffer s & 1 of every CTA must be done with frame s-2 (waiting here, after the arithmetic,
            // gives the signal the length of the butterfly to arrive)
            if (i == 0 && s >= 2) mbar_wait_bounded(bar_free + (s & 1), ((s >> 1) - 1) & 1);
            const uint32_t off = (uint32_t)psg_pad(c * SL + t + i * T) * 8;
#pragma unroll
            for (int k0 = 0; k0 < R0; ++k0) st_async_cf(map_cluster(dst + off, k0), x[k0], map_cluster(fullb, k0));
        }
    };
    auto epilogue = [&](int item, float2* xb) {
        __syncthreads();
        float* sout = reinterpret_cast<float*>(xb);
        const int klow = PL::low_freq(t);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int freq = klow + (N2 / 16) * jj;
            sout[freq ^ (((freq >> 5) & 7) << 2)] = acc[jj];
        }
        __syncthreads();
        const float4* sout4 = reinterpret_cast<const float4*>(sout);
        float4* dst = reinterpret_cast<float4*>(a.tmp + ((size_t)item * R0 + c) * N2);
#pragma unroll
        for (int q = t; q < N2 / 4; q += T) dst[q] = sout4[q ^ ((q >> 3) & 7)];
    };

    __syncthreads();  // steps 0..2 published
    if (ring[0].valid) {
        prepass(0, ring[0].skew);
        __syncthreads();  // the slab stage has been read
        if (t == 0) fetch(1);
        for (int r = 0;; ++r) {
            const int cur_valid = ring[r & 3].valid, cur_last = ring[r & 3].last, cur_item = ring[r & 3].item;
            if (!cur_valid) break;
            const int nxt_valid = ring[(r + 1) & 3].valid, nxt_skew = ring[(r + 1) & 3].skew;
            if (nxt_valid) prepass(r + 1, nxt_skew);
            float2* xb = xb0 + (size_t)(r & 1) * CF::NPAD;
            cf pw[4];  // pass-0 twiddle powers (L1 hits), in flight while the row is awaited
#pragma unroll
            for (int q = 0; q < 4; ++q) pw[q] = __ldg(a.twp + PL::TW0 + ((1 << q) - 1) * PL::S0 + t);
            mbar_wait_bounded(bar_full + (r & 1), (r >> 1) & 1);
            if (t == 0) mbar_expect_tx(bar_full + (r & 1), N2 * 8);  // arm the buffer for frame r+2
            {
                // pass 0 in place: this thread's 16 points are read and written by nobody else
                cf x[E];
                float2* p0 = xb + psg_pad(t);
#pragma unroll
                for (int n = 0; n < 16; ++n) x[n] = p0[pad_off(n * PL::S0)];
                dftR<16>(x);
                twiddle_dfs<16>(x, pw);
#pragma unroll
                for (int k = 0; k < 16; ++k) p0[pad_off(k * PL::S0)] = x[k];
            }
            {
                cf tw1[15];
                rebuild_tw<16>(wb1, tw1);
                __syncthreads();  // pass 0 stored; the slab stage has been read; step r+2 is visible
                if (t == 0) {
                    fetch(r + 2);
                    publish();  // step r+3 (its ring slot held step r-1, read before this barrier)
                }
                smem_pass<E, T, 16, PL::S1, false>(xb, tw1, t, acc);
            }
            __syncwarp();
            smem_pass<E, T, 16, 1, true>(xb, nullptr, t, acc);
            if (cur_last) {
                epilogue(cur_item, xb);
#pragma unroll
                for (int i = 0; i < E; ++i) acc[i] = 0.f;
            }
            // buffer r & 1 is idle: tell every CTA of the cluster (signalling per warp instead, without this
            // barrier, was measured equal)
            __syncthreads();
            if (t < R0) mbar_arrive_cluster(map_cluster(smem_u32(bar_free + (r & 1)), (unsigned)t));
        }
    }
    // peers may still signal this CTA's barriers: leave together
    cluster_arrive();
    cluster_wait();
}

// tmp[cs][split][k0][k'] raw sums -> column cs: sum of the splits (fp64, fixed order), bin
// k = k0 + R0*k' at output index (k + N/2) mod N, scaled, linear and/or dB.  One CTA moves 128
// consecutive k' of one column through shared memory so that both sides are coalesced.
struct ClusterFinArgs {
    const float* tmp;
    int r0, nsplit;
    float scale, eps;
    float* out_lin;
    float* out_db;
};

__global__ void __launch_bounds__(256) sti_cluster_finalize_kernel(const ClusterFinArgs a) {
    constexpr int N2 = 4096, KT = 128;
    __shared__ float tile[16][KT + 1];
    const int r0 = a.r0;
    const int N = r0 * N2;
    const size_t cs = blockIdx.y;
    const int kp0 = blockIdx.x * KT;
    for (int e = threadIdx.x; e < r0 * KT; e += 256) {
        const int k0 = e / KT, j = e - k0 * KT;
        const float* p = a.tmp + ((cs * a.nsplit) * r0 + k0) * N2 + kp0 + j;
        float v;
        if (a.nsplit == 1) {
            v = p[0] * a.scale;
        } else {
            double s = 0.0;
            for (int sp = 0; sp < a.nsplit; ++sp) s += (double)p[(size_t)sp * r0 * N2];
            v = (float)(s * (double)a.scale);
        }
        tile[k0][j] = v;
    }
    __syncthreads();
    const size_t ocol = cs * (size_t)N;
    for (int e = threadIdx.x; e < r0 * KT; e += 256) {
        const int j = e / r0, k0 = e - j * r0;
        const int k = k0 + r0 * (kp0 + j);
        const int idx = (k + N / 2) & (N - 1);
        const float v = tile[k0][j];
        if (a.out_lin) a.out_lin[ocol + idx] = v;
        if (a.out_db) a.out_db[ocol + idx] = power_to_db(v, a.eps);
    }
}
