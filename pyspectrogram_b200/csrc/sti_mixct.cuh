// Round FFT lengths (1000, 1200, 2000, 3000, 4000, 5000, 10000 ...: the numbers people type into the viewer's
// nfft box, drfview.py:474-479) as ONE mixed-radix transform with every stride known at compile time.
//
// sti_mixed_kernel (sti_bluestein.cuh) covers every N = 2^a 3^b 5^c with run-time radices: integer divisions per
// butterfly, twiddles and window loaded per element, accumulators read-modify-written in shared memory -- 11 % of
// the HBM peak at nfft = 1000.  Here a plan is a template: <N, R0..R3, T> with radices up to 20 that may be
// composite -- 6, 10, 12, 15, 20 run as Good-Thomas prime-factor butterflies (two small DFTs, no internal twiddles)
// -- so 1000 = 10 x 10 x 10 takes three passes (five shared-memory accesses per sample) where 8 x 5 x 5 x 5 took four.
//   * a frame group of T threads owns one frame; thread t runs butterflies bf = t + i T of every pass, the same
//     ones for every frame, so its window values, its twiddles W_N^{n' k N / (R_p S_p)} (TWREG plans; the others
//     load W^1, W^2, W^4, W^8 from the L1-resident table and multiply) and its |X|^2 accumulators live in registers;
//   * pass 0 reads the samples with coalesced LDG (any layout and sample type), the next frame's loads are issued
//     before the current frame's last pass; passes run in place on a padded shared-memory buffer (padding chosen
//     per plan by tools/mixct_pad_search.py, linear inside a butterfly: one padded base + immediate offsets); one
//     CTA barrier per pass;
//   * F groups per CTA share the barriers and work on F frames of the same column (the shipped plans use F = 1 and
//     several resident CTAs: more independent barrier domains per SM measured faster); the epilogue sums their
//     accumulators through shared memory, un-permutes the mixed-radix digit reversal, fftshifts (odd N included)
//     and stores 10 log10 / linear power coalesced -- or raw partial sums for split columns (sti_finalize_kernel).
// Pass structure (the same index algebra as sti_kernels.cuh, restated in numpy in tests/test_fft_plan.py):
//   before pass p the element with digits (k_0..k_{p-1}, n_rest) sits at pos = sum_q k_q S_q + n_rest,
//   S_p = N / (R_0..R_p); pass p: n_rest = n_p S_p + n', R_p-point DFT over n_p, output k_p times
//   W_{R_p S_p}^{n' k_p}, stored in place of n_p; after the last pass pos holds frequency sum_q k_q R_0..R_{q-1}.
#pragma once
#include "sti_common.cuh"

// ---- butterflies ----------------------------------------------------------------------------------------------
PSG_HD void mx_dft3(cf* a) {
    const cf t = cadd(a[1], a[2]);
    const cf u = fma2(t, make_float2(-0.5f, -0.5f), a[0]);
    const cf d = cscale(csub(a[1], a[2]), 0.86602540378443864676f);
    a[0] = cadd(a[0], t);
    a[1] = cadd(u, mul_nj(d));
    a[2] = csub(u, mul_nj(d));
}
PSG_HD void mx_dft5(cf* a) {
    constexpr float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    constexpr float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const cf t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]), t3 = csub(a[1], a[4]), t4 = csub(a[2], a[3]);
    const cf m1 = fma2(t2, make_float2(c2, c2), fma2(t1, make_float2(c1, c1), a[0]));
    const cf m2 = fma2(t2, make_float2(c1, c1), fma2(t1, make_float2(c2, c2), a[0]));
    const cf n1 = fma2(t4, make_float2(s2, s2), cscale(t3, s1));
    const cf n2 = fma2(t4, make_float2(-s1, -s1), cscale(t3, s2));
    a[0] = cadd(a[0], cadd(t1, t2));
    a[1] = cadd(m1, mul_nj(n1));
    a[4] = csub(m1, mul_nj(n1));
    a[2] = cadd(m2, mul_nj(n2));
    a[3] = csub(m2, mul_nj(n2));
}
template <int R>
PSG_HD void mx_dft(cf* v);
__host__ __device__ constexpr int mx_inv_mod(int a, int m) {  // a^-1 mod m (small, coprime)
    for (int x = 1; x < m; ++x)
        if ((a * x) % m == 1) return x;
    return 1;
}
// Good-Thomas: R = A B, gcd(A, B) = 1.  n = (B n1 + A n2) mod R, k = (B (B^-1 mod A) k1 + A (A^-1 mod B) k2) mod R
// give W_R^{nk} = W_A^{n1 k1} W_B^{n2 k2}: A-point DFTs over n1, B-point DFTs over n2, no twiddles in between; the
// index maps are compile-time (register renaming).
template <int A, int B>
PSG_HD void mx_dft_pfa(cf* v) {
    constexpr int R = A * B, P = mx_inv_mod(B % A, A), Q = mx_inv_mod(A % B, B);
    cf x[B][A];  // x[n2][n1]
#pragma unroll
    for (int n2 = 0; n2 < B; ++n2)
#pragma unroll
        for (int n1 = 0; n1 < A; ++n1) x[n2][n1] = v[(B * n1 + A * n2) % R];
#pragma unroll
    for (int n2 = 0; n2 < B; ++n2) mx_dft<A>(x[n2]);  // -> x[n2][k1]
#pragma unroll
    for (int k1 = 0; k1 < A; ++k1) {
        cf y[B];
#pragma unroll
        for (int n2 = 0; n2 < B; ++n2) y[n2] = x[n2][k1];
        mx_dft<B>(y);
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) v[(B * P * k1 + A * Q * k2) % R] = y[k2];
    }
}
template <int R>
PSG_HD void mx_dft(cf* v) {
    if constexpr (R == 6) mx_dft_pfa<2, 3>(v);
    else if constexpr (R == 10) mx_dft_pfa<2, 5>(v);
    else if constexpr (R == 12) mx_dft_pfa<4, 3>(v);
    else if constexpr (R == 15) mx_dft_pfa<3, 5>(v);
    else if constexpr (R == 20) mx_dft_pfa<4, 5>(v);
    else if constexpr (R == 3) mx_dft3(v);
    else if constexpr (R == 5) mx_dft5(v);
    else dftR<R>(v);
}

// ---- plan -------------------------------------------------------------------------------------------------------
// Padded address of pos: pos + PA * (pos / PQ) + PA2 * (pos / PQ2)  (a PQ of 0 switches its term off).  PQ is the last
// radix or a multiple of it and PQ2 a product of trailing radices, so that inside one butterfly the padding is LINEAR:
// every stride S_p is a multiple of PQ (and of PQ2, or the whole butterfly lies inside one PQ2 block), and the last
// pass's run of consecutive positions stays inside one block.  A thread therefore pads its base once and reaches its R
// elements with compile-time offsets (off()): LDS / STS with immediate offsets, no address arithmetic per element --
// the first version divided per element, which cost more than the conflicts it removed (profiles/r02_mixct_alternatives.txt).
// TW_: 0 = twiddles rebuilt per frame from W^(2^q) loaded from the L1-resident table, 1 = all twiddles in registers.
// (A third form that kept neither the next frame's samples nor the window in registers -- no spills on the radix-20
// plans -- measured 20 % slower than spilling: the exposed load latency costs more.  Removed.)
template <int N_, int R0_, int R1_, int R2_, int R3_, int T_, int PQ_, int PA_, int TW_, int PQ2_ = 0, int PA2_ = 0>
struct MixPlan {
    static constexpr int N = N_, T = T_, PQ = PQ_, PA = PA_, PQ2 = PQ2_, PA2 = PA2_;
    static constexpr bool TWREG = TW_ == 1;
    static constexpr int P = 2 + (R2_ > 1) + (R3_ > 1);
    static_assert(R0_ * R1_ * R2_ * R3_ == N_ && R1_ > 1 && (R3_ == 1 || R2_ > 1), "radices multiply to N; at least two passes");
    __host__ __device__ static constexpr int r(int p) { return p == 0 ? R0_ : p == 1 ? R1_ : p == 2 ? R2_ : R3_; }
    __host__ __device__ static constexpr int s(int p) {
        return p == 0 ? N_ / R0_ : p == 1 ? N_ / (R0_ * R1_) : p == 2 ? N_ / (R0_ * R1_ * R2_) : N_ / (R0_ * R1_ * R2_ * R3_);
    }
    __host__ __device__ static constexpr int nb(int p) { return (N_ / r(p) + T_ - 1) / T_; }  // butterflies per thread in pass p
    static constexpr int RL = r(P - 1), NBL = nb(P - 1);
    static constexpr int NPADDED = N_ + (PQ_ ? PA_ * ((N_ + PQ_ - 1) / PQ_) : 0) + (PQ2_ ? PA2_ * ((N_ + PQ2_ - 1) / PQ2_) : 0);
    static constexpr int BUF = (NPADDED + 3) & ~3;  // complex per group buffer
    // linearity of the padding inside the butterflies of pass p (see above)
    __host__ __device__ static constexpr bool lin_ok(int p) {
        const int sp = s(p), span = r(p) * s(p);
        const bool t1 = PQ_ == 0 || (p == P - 1 ? PQ_ % r(p) == 0 : sp % PQ_ == 0);
        const bool t2 = PQ2_ == 0 || (p == P - 1 ? PQ2_ % r(p) == 0 : (sp % PQ2_ == 0 || PQ2_ % span == 0));
        return t1 && t2;
    }
    static_assert(lin_ok(0) && lin_ok(1) && (P < 3 || lin_ok(2)) && (P < 4 || lin_ok(3)), "padding must be linear inside a butterfly");
    // padded offset of element n of a butterfly of pass p from the padded base
    __host__ __device__ static constexpr int off(int p, int n) {
        const int d = n * s(p);
        return d + (PQ_ && d % PQ_ == 0 ? PA_ * (d / PQ_) : 0) + (PQ2_ && d % PQ2_ == 0 ? PA2_ * (d / PQ2_) : 0);
    }
    // twiddle registers of a TWREG plan (complex values): passes 0 .. P-2
    __host__ __device__ static constexpr int tw_off(int p) {
        int c = 0;
        for (int q = 0; q < p; ++q) c += nb(q) * (r(q) - 1);
        return c;
    }
    static constexpr int NTW = tw_off(P - 1);
    PSG_HD static int pad(int pos) {
        int a = pos;
        if constexpr (PQ_ != 0) a += PA_ * (pos / PQ_);
        if constexpr (PQ2_ != 0) a += PA2_ * (pos / PQ2_);
        return a;
    }
    // frequency held by position pos after the last pass
    PSG_HD static int freq(int pos) {
        int f = 0, mul = 1, rem = pos;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int d = rem / s(p);
            rem -= d * s(p);
            f += d * mul;
            mul *= r(p);
        }
        return f;
    }
};

// W^k, k = 1 .. R-1, from the powers pw[q] = W^(2^q): one complex multiply per k that is not a power of two
__host__ __device__ constexpr int mx_log2(int k) { return k >= 16 ? 4 : k >= 8 ? 3 : k >= 4 ? 2 : k >= 2 ? 1 : 0; }
__host__ __device__ constexpr int mx_npow(int r) { return mx_log2(r - 1) + 1; }  // powers needed for k <= r - 1
template <int R>
PSG_HD void mx_twiddle_from_powers(cf* v, const cf* pw) {
    cf w[R];
#pragma unroll
    for (int k = 1; k < R; ++k) {
        if ((k & (k - 1)) == 0) {
            w[k] = pw[mx_log2(k)];
        } else {
            const int low = k & (-k);
            w[k] = cmul(w[k - low], w[low]);
        }
        v[k] = cmul(v[k], w[k]);
    }
}

// pass PIDX >= 1 of one frame, in place; the last pass adds |X|^2 to the thread's accumulators
template <class PL, int PIDX>
PSG_DEV void mx_pass(float2* __restrict__ buf, int t, const cf* twr, const float2* __restrict__ twf, float* acc) {
    constexpr int N = PL::N, T = PL::T, RR = PL::r(PIDX), Sp = PL::s(PIDX), NBP = PL::nb(PIDX);
    constexpr bool LAST = PIDX == PL::P - 1;
#pragma unroll
    for (int i = 0; i < NBP; ++i) {
        const int bf = t + i * T;
        if (NBP * T == N / RR || bf < N / RR) {
            const int blk = bf / Sp, npr = bf - blk * Sp;
            const int base = blk * RR * Sp + npr;
            float2* const pb = buf + PL::pad(base);
            cf v[RR];
#pragma unroll
            for (int n = 0; n < RR; ++n) v[n] = pb[PL::off(PIDX, n)];
            mx_dft<RR>(v);
            if constexpr (LAST) {
#pragma unroll
                for (int k = 0; k < RR; ++k) acc[i * RR + k] = fmaf(v[k].x, v[k].x, fmaf(v[k].y, v[k].y, acc[i * RR + k]));
            } else {
                if constexpr (PL::TWREG) {
#pragma unroll
                    for (int k = 1; k < RR; ++k) v[k] = cmul(v[k], twr[PL::tw_off(PIDX) + i * (RR - 1) + k - 1]);
                } else {
                    constexpr int ts = N / (RR * Sp), NPW = mx_npow(RR);
                    cf pw[NPW];
#pragma unroll
                    for (int q = 0; q < NPW; ++q) pw[q] = __ldg(twf + (((npr * ts) << q) % N));
                    mx_twiddle_from_powers<RR>(v, pw);
                }
#pragma unroll
                for (int k = 0; k < RR; ++k) pb[PL::off(PIDX, k)] = v[k];
            }
        }
    }
}

// twiddle registers of pass PIDX (TWREG plans)
template <class PL, int PIDX>
PSG_DEV void mx_load_tw(cf* twr, int t, const float2* __restrict__ twf) {
    constexpr int N = PL::N, T = PL::T, RR = PL::r(PIDX), Sp = PL::s(PIDX), ts = N / (RR * Sp);
#pragma unroll
    for (int i = 0; i < PL::nb(PIDX); ++i) {
        const int bf = t + i * T;
        const int npr = bf % Sp;
#pragma unroll
        for (int k = 1; k < RR; ++k)
            twr[PL::tw_off(PIDX) + i * (RR - 1) + k - 1] = (bf < N / RR) ? __ldg(twf + npr * k * ts) : make_float2(1.f, 0.f);
    }
}

template <class PL, int F, int IQT, int MINB>
__global__ void __launch_bounds__(F * PL::T, MINB) sti_mixct_kernel(const StiArgs a) {
    constexpr int N = PL::N, T = PL::T, P = PL::P, NT = F * T;
    constexpr int R0 = PL::r(0), NB0 = PL::nb(0), S0 = PL::s(0);
    constexpr int RL = PL::RL, NBL = PL::NBL;
    extern __shared__ __align__(16) float2 mx_smem[];
    const int g = threadIdx.x / T, t = threadIdx.x - g * T;
    float2* const buf = mx_smem + (size_t)g * PL::BUF;

    // ---- loop-invariant per-thread tables: window, twiddles ----
    float win[NB0 * R0];
#pragma unroll
    for (int i = 0; i < NB0; ++i)
#pragma unroll
        for (int n = 0; n < R0; ++n) {
            const int bf = t + i * T;
            win[i * R0 + n] = (bf < S0) ? __ldg(a.win + bf + n * S0) : 0.f;
        }
    cf twr[PL::TWREG ? PL::NTW : 1];
    if constexpr (PL::TWREG) {
        mx_load_tw<PL, 0>(twr, t, a.tw);
        if constexpr (P > 2) mx_load_tw<PL, 1>(twr, t, a.tw);
        if constexpr (P > 3) mx_load_tw<PL, 2>(twr, t, a.tw);
    }

    const int ncs = a.ncol * a.nsub;
    for (int item = blockIdx.x; item < ncs * a.nsplit; item += gridDim.x) {
        const int split = item % a.nsplit, cs = item / a.nsplit;
        const int col = cs % a.ncol, sub = cs / a.ncol;
        const int k0 = split * a.chunk, k1 = min(a.nfr, k0 + a.chunk);
        const long long src0 = __ldg(a.col_off + col) + (long long)sub * a.sub_stride;
        float acc[NBL * RL];
#pragma unroll
        for (int i = 0; i < NBL * RL; ++i) acc[i] = 0.f;
        const int niter = (k1 - k0 + F - 1) / F;
        // samples of this thread's pass-0 butterflies of frame k (zeros for a group without a frame)
        cf nx[NB0 * R0];
        auto load_frame = [&](int k) {
            const bool live = k < k1;
            const long long src = src0 + (long long)k * a.hop_elems;
#pragma unroll
            for (int i = 0; i < NB0; ++i)
#pragma unroll
                for (int n = 0; n < R0; ++n) {
                    const int bf = t + i * T;
                    nx[i * R0 + n] = (live && bf < S0) ? ldg_iq<IQT>(a.iq, src + (long long)(bf + n * S0) * a.sample_stride)
                                                        : make_float2(0.f, 0.f);
                }
        };
        load_frame(k0 + g);
        for (int j = 0; j < niter; ++j) {
            // ---- pass 0: window, R0-point DFT, twiddle, store ----
            cf x[NB0 * R0];
#pragma unroll
            for (int i = 0; i < NB0 * R0; ++i) x[i] = mul2(nx[i], make_float2(win[i], win[i]));
            if constexpr (P == 2) load_frame(k0 + (j + 1) * F + g);
            __syncthreads();  // the previous frame's last pass is done reading the buffer
#pragma unroll
            for (int i = 0; i < NB0; ++i) {
                const int bf = t + i * T;
                mx_dft<R0>(&x[i * R0]);
                if (NB0 * T == S0 || bf < S0) {
                    if constexpr (PL::TWREG) {
#pragma unroll
                        for (int k = 1; k < R0; ++k) x[i * R0 + k] = cmul(x[i * R0 + k], twr[i * (R0 - 1) + k - 1]);
                    } else {
                        constexpr int NPW = mx_npow(R0);
                        cf pw[NPW];
#pragma unroll
                        for (int q = 0; q < NPW; ++q) pw[q] = __ldg(a.tw + ((bf << q) % N));  // R0 S0 = N: W_N^{n' 2^q}
                        mx_twiddle_from_powers<R0>(&x[i * R0], pw);
                    }
                    float2* const pb = buf + PL::pad(bf);
#pragma unroll
                    for (int k = 0; k < R0; ++k) pb[PL::off(0, k)] = x[i * R0 + k];
                }
            }
            // ---- passes 1 .. P-1 in place; the next frame's loads go out before the last one ----
            __syncthreads();
            if constexpr (P == 2) {
                mx_pass<PL, 1>(buf, t, twr, a.tw, acc);
            } else {
                mx_pass<PL, 1>(buf, t, twr, a.tw, acc);
                __syncthreads();
                if constexpr (P == 3) {
                    load_frame(k0 + (j + 1) * F + g);
                    mx_pass<PL, 2>(buf, t, twr, a.tw, acc);
                } else {
                    mx_pass<PL, 2>(buf, t, twr, a.tw, acc);
                    __syncthreads();
                    load_frame(k0 + (j + 1) * F + g);
                    mx_pass<PL, 3>(buf, t, twr, a.tw, acc);
                }
            }
        }
        // ---- epilogue: group sums, digit reversal, fftshift, scale, store ----
        __syncthreads();  // every group is done with its buffer
        float* const sout = reinterpret_cast<float*>(mx_smem);  // [F][2 BUF] floats: a group's buffer holds >= N of them
        constexpr int half = N / 2;  // np.fft.fftshift: out[(k + N//2) mod N] = in[k]
#pragma unroll
        for (int i = 0; i < NBL; ++i) {
            const int bf = t + i * T;
            if (NBL * T == N / RL || bf < N / RL) {
#pragma unroll
                for (int k = 0; k < RL; ++k) {
                    int idx = PL::freq(bf * RL + k) + half;
                    if (idx >= N) idx -= N;
                    sout[(size_t)g * (2 * PL::BUF) + idx] = acc[i * RL + k];
                }
            }
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < N; idx += NT) {
            float sum = sout[idx];
#pragma unroll
            for (int gg = 1; gg < F; ++gg) sum += sout[(size_t)gg * (2 * PL::BUF) + idx];
            if (a.nsplit > 1) {
                a.partial[((size_t)cs * a.nsplit + split) * N + idx] = sum;
            } else {
                const float pw = sum * a.scale;
                if (a.out_lin) a.out_lin[(size_t)cs * N + idx] = pw;
                if (a.out_db) a.out_db[(size_t)cs * N + idx] = power_to_db(pw, a.eps);
            }
        }
        // (the next item's first store into the buffers comes after a CTA barrier)
    }
}
