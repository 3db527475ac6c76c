// Arbitrary (non power-of-two) nfft, Bluestein's chirp-z algorithm with high-radix passes.
//
// Same mathematics as sti_bluestein_kernel (sti_kernels.cuh): with c[n] = exp(+j*pi*n^2/N),
//   |X[k]|^2 = |(a (*) c)[k]|^2,  a[n] = x[n] w[n] conj(c[n]),
// the circular convolution of length M = 2^m >= 2N-1 computed as IFFT_M(FFT_M(a) .* B), B = FFT_M(c)/M.
// That kernel runs 2*log2(M) radix-2 passes over shared memory per frame (22 exchanges at M = 2048);
// this one runs the M-point transforms as mixed-radix passes (first radix 2, 4 or 8, then 16s -- the
// butterflies of cplx.cuh) in the padded in-place layout of the tuned kernels:
//   forward   pass p splits n_rest = n_p*S_p + n', R_p-point DFT over n_p, output k_p times
//             W_M^{n' k_p M/(R_p S_p)}, stored in place of n_p (the index algebra of sti_kernels.cuh).  Pass 0
//             reads the samples straight from global memory (zero beyond N) and applies a[n]; the last pass
//             multiplies by B, which the host stores in the order the positions come out (digit-reversed).
//   inverse   the transposed network, passes P-1 .. 0: conjugate twiddle, then the inverse R_p-point DFT
//             (conj . DFT . conj) over k_p; positions come back in natural order and the last pass adds
//             |y[n]|^2, n < N, to the column's accumulators in shared memory instead of storing y.
// 2P-1 exchanges per frame (5 at M = 2048, 7 at M = 16384).  A CTA is G frame groups of max(16, M/16)
// threads working on G frames of one work item (column, frame chunk) at a time; the radix-2 kernel remains
// the path for M > 16384 (global scratch).
#pragma once
#include "sti_cluster.cuh"

struct Bluestein16Args {
    const float2* aw;    // [N]  w[n]/sum(w) * conj(c[n])
    const float2* bpos;  // [M]  FFT_M(c)[freq(pos)] / M, indexed by position after the forward passes
    const float2* twf;   // [M]  exp(-2*pi*j*m/M)
    int n, logm;
    int tpf;  // threads per frame group (divides the CTA size)
    int npass;
    int radix[4];  // radix[0] in {2, 4, 8, 16}, the rest 16
};

PSG_DEV cf cconj(cf a) { return make_float2(a.x, -a.y); }

// forward pass: FIRST reads global samples * aw, LAST multiplies by B
template <int R>
PSG_DEV void bs_forward_pass(float2* __restrict__ buf, int logS, int M, int Mv, int tid, int nt, const float2* __restrict__ twf,
                             bool first, bool last, const StiArgs& a, long long src, const Bluestein16Args& b) {
    const int S = 1 << logS;
    constexpr int NPW = psg_npow(R);
    const int tstride = M / (R << logS);  // W_{R S}^{e} = W_M^{e * tstride}
    for (int bf = tid; bf < Mv / R; bf += nt) {
        const int npr = bf & (S - 1);
        const int base = ((bf >> logS) * R << logS) + npr;
        cf v[R];
        if (first) {
#pragma unroll
            for (int n = 0; n < R; ++n) {
                const int i = base + (n << logS);
                v[n] = (i < b.n) ? cmul(ldg_iq_rt(a.iq_type, a.iq, src + (long long)i * a.sample_stride), __ldg(b.aw + i))
                                 : make_float2(0.f, 0.f);
            }
        } else {
#pragma unroll
            for (int n = 0; n < R; ++n) v[n] = buf[psg_pad(base + (n << logS))];
        }
        dftR<R>(v);
        if (last) {
#pragma unroll
            for (int k = 0; k < R; ++k) v[k] = cmul(v[k], __ldg(b.bpos + base + (k << logS)));
        } else {
            // W^k, k < R, from W^1, W^2, W^4, W^8 (four loads instead of fifteen; sti_cluster.cuh)
            cf pw[NPW];
#pragma unroll
            for (int q = 0; q < NPW; ++q) pw[q] = __ldg(twf + ((npr * tstride) << q));
            twiddle_dfs<R>(v, pw);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) buf[psg_pad(base + (k << logS))] = v[k];
    }
}

// inverse of the forward pass with the same (R, S): FIRST here is the pass that has no twiddle (the
// forward network's last pass), FINAL accumulates |y|^2 for n < nvalid instead of storing
template <int R>
PSG_DEV void bs_inverse_pass(float2* __restrict__ buf, int logS, int M, int Mv, int tid, int nt, const float2* __restrict__ twf,
                             bool notw, bool final_pass, float* __restrict__ accs, int nvalid) {
    const int S = 1 << logS;
    constexpr int NPW = psg_npow(R);
    const int tstride = M / (R << logS);
    for (int bf = tid; bf < Mv / R; bf += nt) {
        const int npr = bf & (S - 1);
        const int base = ((bf >> logS) * R << logS) + npr;
        cf v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = buf[psg_pad(base + (k << logS))];
        if (!notw) {
            cf pw[NPW];
#pragma unroll
            for (int q = 0; q < NPW; ++q) pw[q] = cconj(__ldg(twf + ((npr * tstride) << q)));
            twiddle_dfs<R>(v, pw);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = cconj(v[k]);
        dftR<R>(v);  // conj(IDFT) of the inputs: the conjugate is undone below or irrelevant for |.|^2
        if (final_pass) {
#pragma unroll
            for (int n = 0; n < R; ++n) {
                const int i = base + (n << logS);
                if (i < nvalid) accs[i] = fmaf(v[n].x, v[n].x, fmaf(v[n].y, v[n].y, accs[i]));
            }
        } else {
#pragma unroll
            for (int n = 0; n < R; ++n) buf[psg_pad(base + (n << logS))] = cconj(v[n]);
        }
    }
}

PSG_DEV void bs_forward_dispatch(int R, float2* buf, int logS, int M, int Mv, int tid, int nt, const float2* twf, bool first,
                                 bool last, const StiArgs& a, long long src, const Bluestein16Args& b) {
    switch (R) {
        case 2: bs_forward_pass<2>(buf, logS, M, Mv, tid, nt, twf, first, last, a, src, b); break;
        case 4: bs_forward_pass<4>(buf, logS, M, Mv, tid, nt, twf, first, last, a, src, b); break;
        case 8: bs_forward_pass<8>(buf, logS, M, Mv, tid, nt, twf, first, last, a, src, b); break;
        default: bs_forward_pass<16>(buf, logS, M, Mv, tid, nt, twf, first, last, a, src, b); break;
    }
}
PSG_DEV void bs_inverse_dispatch(int R, float2* buf, int logS, int M, int Mv, int tid, int nt, const float2* twf, bool notw,
                                 bool final_pass, float* accs, int nvalid) {
    switch (R) {
        case 2: bs_inverse_pass<2>(buf, logS, M, Mv, tid, nt, twf, notw, final_pass, accs, nvalid); break;
        case 4: bs_inverse_pass<4>(buf, logS, M, Mv, tid, nt, twf, notw, final_pass, accs, nvalid); break;
        case 8: bs_inverse_pass<8>(buf, logS, M, Mv, tid, nt, twf, notw, final_pass, accs, nvalid); break;
        default: bs_inverse_pass<16>(buf, logS, M, Mv, tid, nt, twf, notw, final_pass, accs, nvalid); break;
    }
}

__global__ void __launch_bounds__(512) sti_bluestein16_kernel(const StiArgs a, const Bluestein16Args b) {
    const int N = b.n, M = 1 << b.logm;
    extern __shared__ __align__(16) float2 bs_smem[];
    // frame groups of b.tpf threads: group g transforms frames k0 + g, k0 + g + G, ... of the item in its own
    // buffer and adds into its own accumulators; the groups are summed in fixed order in the epilogue
    const int T = b.tpf, G = blockDim.x / T;
    const int g = threadIdx.x / T, t = threadIdx.x - g * T;
    const int bstride = psg_pad(M) + 2;          // complex per group buffer (even: 16-byte aligned)
    const int astride = (N + 3) & ~3;            // floats per group accumulator
    float2* buf = bs_smem + (size_t)g * bstride;
    float* acc0 = reinterpret_cast<float*>(bs_smem + (size_t)G * bstride);
    float* accs = acc0 + (size_t)g * astride;
    const int ncs = a.ncol * a.nsub;
    const int P = b.npass;
    for (int item = blockIdx.x; item < ncs * a.nsplit; item += gridDim.x) {
        const int split = item % a.nsplit;
        const int cs = item / a.nsplit;
        const int col = cs % a.ncol, sub = cs / a.ncol;
        const int k0 = split * a.chunk;
        const int k1 = min(a.nfr, k0 + a.chunk);
        const long long src0 = a.col_off[col] + (long long)sub * a.sub_stride;
        __syncthreads();  // the previous item's epilogue is done with the accumulators
        for (int i = t; i < N; i += T) accs[i] = 0.f;
        const int niter = (k1 - k0 + G - 1) / G;
        for (int j = 0; j < niter; ++j) {
            const int k = k0 + j * G + g;
            const int Mv = (k < k1) ? M : 0;  // a group without a frame runs the barriers only
            const long long src = src0 + (long long)k * a.hop_elems;
            int logS = b.logm;
            for (int p = 0; p < P; ++p) {
                const int R = b.radix[p];
                logS -= (R == 16) ? 4 : (R == 8) ? 3 : (R == 4) ? 2 : 1;
                __syncthreads();  // p = 0: the previous frame's last inverse pass has read buf (and accs are zeroed)
                bs_forward_dispatch(R, buf, logS, M, Mv, t, T, b.twf, p == 0, p == P - 1, a, src, b);
            }
            // logS == 0 here; walk the strides back up
            for (int p = P - 1; p >= 0; --p) {
                const int R = b.radix[p];
                __syncthreads();
                bs_inverse_dispatch(R, buf, logS, M, Mv, t, T, b.twf, p == P - 1, p == 0, accs, N);
                logS += (R == 16) ? 4 : (R == 8) ? 3 : (R == 4) ? 2 : 1;
            }
        }
        __syncthreads();
        const int half = N / 2;  // np.fft.fftshift: out[(k + N//2) mod N] = in[k]
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            float sum = acc0[i];
            for (int gg = 1; gg < G; ++gg) sum += acc0[(size_t)gg * astride + i];
            int idx = i + half;
            if (idx >= N) idx -= N;
            if (a.nsplit > 1) {
                a.partial[((size_t)cs * a.nsplit + split) * N + idx] = sum;
            } else {
                const float pw = sum * a.scale;
                if (a.out_lin) a.out_lin[(size_t)cs * N + idx] = pw;
                if (a.out_db) a.out_db[(size_t)cs * N + idx] = power_to_db(pw, a.eps);
            }
        }
    }
}

// ---- nfft = 2^a 3^b 5^c (1000, 1200, 3000, 5000, 10000, ...): direct mixed-radix transform ------------------
// The lengths people type into the viewer's spin box are mostly "round" numbers.  They need no chirp-z
// detour: one N-point transform with radices 16/8/4/2, 5 and 3 in the same in-place pass structure
//   pass p: n_rest = n_p S_p + n', R_p-point DFT over n_p, output k_p times W_N^{n' k_p N/(R_p S_p)}, stored in
//   place of n_p; after the last pass position sum_q k_q S_q holds frequency k_0 + R_0 k_1 + R_0 R_1 k_2 + ...
// instead of two M-point transforms with M >= 2N-1 (a sixth of the butterflies at N = 1000).  Pass 0 reads
// the samples from global memory and applies the window; the last pass adds |X|^2 to accumulators indexed by
// position, and the epilogue un-permutes (mixed-radix digit reversal) and fftshifts once per work item.
struct MixedArgs {
    const float2* twf;  // [N] exp(-2*pi*j*m/N)
    int n, tpf;
    int npass;
    int radix[10];
};

PSG_DEV void dft3(cf* a) {
    const cf t = cadd(a[1], a[2]);
    const cf u = fma2(t, make_float2(-0.5f, -0.5f), a[0]);
    const cf d = cscale(csub(a[1], a[2]), 0.86602540378443864676f);
    a[0] = cadd(a[0], t);
    a[1] = cadd(u, mul_nj(d));
    a[2] = csub(u, mul_nj(d));
}
PSG_DEV void dft5(cf* a) {
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
// Odd prime radices 7, 11, 13 (round 2: lengths such as 1001, 1400, 7000, 9100 leave the Bluestein kernels):
// X[k] = a0 + sum_j (a_j + a_{P-j}) cos(2 pi j k / P) - i sum_j (a_j - a_{P-j}) sin(2 pi j k / P), X[P-k] its mirror:
// (P-1)^2 / 2 packed multiply-adds, no recursion.  cos / sin tables indexed by (j k) mod P fold to immediates.
template <int P>
struct PrimeTab;
template <>
struct PrimeTab<7> {
    static __device__ __forceinline__ float c(int m) {
        constexpr float t[7] = {1.0f, 0.623489802f, -0.222520934f, -0.900968868f, -0.900968868f, -0.222520934f, 0.623489802f};
        return t[m];
    }
    static __device__ __forceinline__ float s(int m) {
        constexpr float t[7] = {0.0f, 0.781831482f, 0.974927912f, 0.433883739f, -0.433883739f, -0.974927912f, -0.781831482f};
        return t[m];
    }
};
template <>
struct PrimeTab<11> {
    static __device__ __forceinline__ float c(int m) {
        constexpr float t[11] = {1.0f, 0.841253533f, 0.415415013f, -0.142314838f, -0.654860734f, -0.959492974f, -0.959492974f, -0.654860734f, -0.142314838f, 0.415415013f, 0.841253533f};
        return t[m];
    }
    static __device__ __forceinline__ float s(int m) {
        constexpr float t[11] = {0.0f, 0.540640817f, 0.909631995f, 0.989821442f, 0.755749574f, 0.281732557f, -0.281732557f, -0.755749574f, -0.989821442f, -0.909631995f, -0.540640817f};
        return t[m];
    }
};
template <>
struct PrimeTab<13> {
    static __device__ __forceinline__ float c(int m) {
        constexpr float t[13] = {1.0f, 0.885456026f, 0.568064747f, 0.12053668f, -0.354604887f, -0.748510748f, -0.970941817f, -0.970941817f, -0.748510748f, -0.354604887f, 0.12053668f, 0.568064747f, 0.885456026f};
        return t[m];
    }
    static __device__ __forceinline__ float s(int m) {
        constexpr float t[13] = {0.0f, 0.464723172f, 0.822983866f, 0.992708874f, 0.935016243f, 0.663122658f, 0.239315664f, -0.239315664f, -0.663122658f, -0.935016243f, -0.992708874f, -0.822983866f, -0.464723172f};
        return t[m];
    }
};
template <int P>
PSG_DEV void dft_oddprime(cf* a) {
    constexpr int H = (P - 1) / 2;
    cf sm[H], df[H];
#pragma unroll
    for (int j = 1; j <= H; ++j) {
        sm[j - 1] = cadd(a[j], a[P - j]);
        df[j - 1] = csub(a[j], a[P - j]);
    }
    cf x0 = a[0];
#pragma unroll
    for (int j = 0; j < H; ++j) x0 = cadd(x0, sm[j]);
    const cf a0 = a[0];
    a[0] = x0;
#pragma unroll
    for (int k = 1; k <= H; ++k) {
        cf A = a0, B = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 1; j <= H; ++j) {
            const float c = PrimeTab<P>::c((j * k) % P), s = PrimeTab<P>::s((j * k) % P);
            A = fma2(sm[j - 1], make_float2(c, c), A);
            B = fma2(df[j - 1], make_float2(s, s), B);
        }
        a[k] = cadd(A, mul_nj(B));      // A - i B
        a[P - k] = csub(A, mul_nj(B));  // A + i B
    }
}
template <int R>
PSG_DEV void dft_any(cf* v) {
    if constexpr (R == 3) dft3(v);
    else if constexpr (R == 5) dft5(v);
    else if constexpr (R == 7 || R == 11 || R == 13) dft_oddprime<R>(v);
    else dftR<R>(v);
}

template <int R>
PSG_DEV void mixed_pass(float2* __restrict__ buf, int S, int N, int Nv, int tid, int nt, const float2* __restrict__ twf,
                        bool first, bool last, const StiArgs& a, long long src, float* __restrict__ accs) {
    const int tstride = N / (R * S);  // W_{R S}^{e} = W_N^{e * tstride}
    for (int bf = tid; bf < Nv / R; bf += nt) {
        const int blk = bf / S;
        const int npr = bf - blk * S;
        const int base = blk * R * S + npr;
        cf v[R];
        if (first) {
#pragma unroll
            for (int n = 0; n < R; ++n) {
                const int i = base + n * S;
                v[n] = cscale(ldg_iq_rt(a.iq_type, a.iq, src + (long long)i * a.sample_stride), __ldg(a.win + i));
            }
        } else {
#pragma unroll
            for (int n = 0; n < R; ++n) v[n] = buf[psg_pad(base + n * S)];
        }
        dft_any<R>(v);
        if (last) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int i = base + k * S;
                accs[i] = fmaf(v[k].x, v[k].x, fmaf(v[k].y, v[k].y, accs[i]));
            }
        } else {
            if constexpr (R == 16 || R == 8) {
                constexpr int NPW = psg_npow(R);
                cf pw[NPW];
#pragma unroll
                for (int q = 0; q < NPW; ++q) pw[q] = __ldg(twf + ((npr * tstride) << q));
                twiddle_dfs<R>(v, pw);
            } else {
#pragma unroll
                for (int k = 1; k < R; ++k) v[k] = cmul(v[k], __ldg(twf + npr * k * tstride));
            }
#pragma unroll
            for (int k = 0; k < R; ++k) buf[psg_pad(base + k * S)] = v[k];
        }
    }
}

__global__ void __launch_bounds__(512) sti_mixed_kernel(const StiArgs a, const MixedArgs b) {
    const int N = b.n;
    extern __shared__ __align__(16) float2 bs_smem[];
    const int T = b.tpf, G = blockDim.x / T;
    const int g = threadIdx.x / T, t = threadIdx.x - g * T;
    const int bstride = (psg_pad(N) + 3) & ~1;  // complex per group buffer (even: 16-byte aligned)
    const int astride = (N + 3) & ~3;           // floats per group accumulator
    float2* buf = bs_smem + (size_t)g * bstride;
    float* acc0 = reinterpret_cast<float*>(bs_smem + (size_t)G * bstride);
    float* accs = acc0 + (size_t)g * astride;
    const int ncs = a.ncol * a.nsub;
    const int P = b.npass;
    for (int item = blockIdx.x; item < ncs * a.nsplit; item += gridDim.x) {
        const int split = item % a.nsplit;
        const int cs = item / a.nsplit;
        const int col = cs % a.ncol, sub = cs / a.ncol;
        const int k0 = split * a.chunk;
        const int k1 = min(a.nfr, k0 + a.chunk);
        const long long src0 = a.col_off[col] + (long long)sub * a.sub_stride;
        __syncthreads();  // the previous item's epilogue is done with the accumulators
        for (int i = t; i < N; i += T) accs[i] = 0.f;
        const int niter = (k1 - k0 + G - 1) / G;
        for (int j = 0; j < niter; ++j) {
            const int k = k0 + j * G + g;
            const int Nv = (k < k1) ? N : 0;  // a group without a frame runs the barriers only
            const long long src = src0 + (long long)k * a.hop_elems;
            int S = N;
            for (int p = 0; p < P; ++p) {
                const int R = b.radix[p];
                S /= R;
                __syncthreads();
                const bool first = p == 0, last = p == P - 1;
                switch (R) {
                    case 2: mixed_pass<2>(buf, S, N, Nv, t, T, b.twf, first, last, a, src, accs); break;
                    case 3: mixed_pass<3>(buf, S, N, Nv, t, T, b.twf, first, last, a, src, accs); break;
                    case 4: mixed_pass<4>(buf, S, N, Nv, t, T, b.twf, first, last, a, src, accs); break;
                    case 5: mixed_pass<5>(buf, S, N, Nv, t, T, b.twf, first, last, a, src, accs); break;
                    case 7: mixed_pass<7>(buf, S, N, Nv, t, T, b.twf, first, last, a, src, accs); break;
                    case 11: mixed_pass<11>(buf, S, N, Nv, t, T, b.twf, first, last, a, src, accs); break;
                    case 13: mixed_pass<13>(buf, S, N, Nv, t, T, b.twf, first, last, a, src, accs); break;
                    case 8: mixed_pass<8>(buf, S, N, Nv, t, T, b.twf, first, last, a, src, accs); break;
                    default: mixed_pass<16>(buf, S, N, Nv, t, T, b.twf, first, last, a, src, accs); break;
                }
            }
        }
        __syncthreads();
        const int half = N / 2;  // np.fft.fftshift: out[(k + N//2) mod N] = in[k]
        for (int pos = threadIdx.x; pos < N; pos += blockDim.x) {
            float sum = acc0[pos];
            for (int gg = 1; gg < G; ++gg) sum += acc0[(size_t)gg * astride + pos];
            // frequency held by this position: digits k_q = (pos / S_q) mod R_q, weight R_0 .. R_{q-1}
            int rem = pos, s = N, mul = 1, freq = 0;
            for (int p = 0; p < P; ++p) {
                const int R = b.radix[p];
                s /= R;
                const int d = rem / s;
                rem -= d * s;
                freq += d * mul;
                mul *= R;
            }
            int idx = freq + half;
            if (idx >= N) idx -= N;
            if (a.nsplit > 1) {
                a.partial[((size_t)cs * a.nsplit + split) * N + idx] = sum;
            } else {
                const float pw = sum * a.scale;
                if (a.out_lin) a.out_lin[(size_t)cs * N + idx] = pw;
                if (a.out_db) a.out_db[(size_t)cs * N + idx] = power_to_db(pw, a.eps);
            }
        }
    }
}
