// Fused STFT -> PSD -> STI kernels for sm_100a (device code).
//
// One CTA owns one work item = (sub-channel, STI column, chunk of that column's frames).  For
// every frame it loads nfft complex64 samples straight into registers (coalesced LDG.64, L1
// no-allocate), multiplies by the fp32 window table w/sum(w), runs an in-place mixed-radix DIF
// FFT whose butterflies live in registers (radix <= 16, packed f32x2 math, cplx.cuh) and whose
// digit exchanges go through a padded, bank-conflict-free shared-memory buffer, and accumulates
// |X|^2 per output bin in registers.  After the chunk's last frame the sums are scaled,
// fftshifted through shared memory and stored coalesced as linear power and/or 10*log10(p+eps)
// -- or as raw partial sums when a column is split over several CTAs (finalised by
// sti_finalize_kernel).  Samples are read from HBM exactly once; nothing but the STI column is
// written.
//
// Index algebra (validated numerically by tests/test_fft_plan.py, which restates it in numpy):
//   N = R0*R1*...*R(P-1),  S_p = N/(R0..Rp).  Before pass p the element with digits
//   (k_0..k_{p-1}, n_rest) sits at pos = sum_q k_q*S_q + n_rest.  Pass p splits
//   n_rest = n_p*S_p + n', does the R_p-point DFT over n_p, multiplies output k_p by
//   W_{R_p*S_p}^{n'*k_p} (skipped on the last pass) and stores it in place of n_p.  After the
//   last pass pos = sum_q k_q*S_q holds frequency k = k_0 + R0*k_1 + R0*R1*k_2 + ...
//   Shared-memory address of pos is pos + (pos >> 4) (one complex of padding per 16), which
//   makes every pass's 64-bit accesses conflict-free for the plans used here.
#pragma once
#include <stdint.h>
#include "cplx.cuh"

struct StiArgs {
    const float2* iq;
    long long sample_stride;  // elements between consecutive samples
    long long sub_stride;     // elements between sub-channels
    long long hop_elems;      // hop * sample_stride
    const long long* col_off; // [ncol] element offset of each column's first sample
    int ncol, nsub;
    int nfr;     // frames per column
    int chunk;   // frames per work item
    int nsplit;  // work items per column = ceil(nfr / chunk)
    const float* win;   // [N]  w[n] / sum(w)
    const float2* tw;   // [N]  exp(-2*pi*j*m/N)
    float scale;        // in_scale^2 / nfr
    float eps;
    float* out_lin;     // [nsub][ncol][N] or null
    float* out_db;      // [nsub][ncol][N] or null
    float* partial;     // [nsub][ncol][nsplit][N] raw sums when nsplit > 1
};

PSG_DEV float2 ldg_stream(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

PSG_DEV float power_to_db(float p, float eps) { return 10.0f * log10f(p + eps); }

__host__ __device__ constexpr int psg_pad(int pos) { return pos + (pos >> 4); }

template <int N, int R0, int R1, int R2, int R3>
struct Plan {
    static constexpr int P = 1 + (R1 > 1) + (R2 > 1) + (R3 > 1);
    static constexpr int S0 = N / R0, S1 = S0 / R1, S2 = S1 / R2, S3 = S2 / R3;
    static_assert(R0 * R1 * R2 * R3 == N, "radices must multiply to N");
    static constexpr int RL = (P == 1) ? R0 : (P == 2) ? R1 : (P == 3) ? R2 : R3;  // last radix
    // frequency of (last-pass butterfly b, output j): digit-reverse b, add (N/RL)*j
    PSG_DEV static int low_freq(int b) {
        int rem = b, k = 0;
        if (P >= 2) { constexpr int s = S0 / RL; k += (rem / s); rem %= s; }
        if (P >= 3) { constexpr int s = S1 / RL; k += (rem / s) * R0; rem %= s; }
        if (P >= 4) { constexpr int s = S2 / RL; k += (rem / s) * R0 * R1; rem %= s; }
        return k;
    }
};

// one in-place pass over shared memory (p >= 1). LAST: accumulate |X|^2 instead of storing.
template <int N, int E, int T, int R, int S, bool LAST>
PSG_DEV void smem_pass(float2* __restrict__ buf, const float2* __restrict__ tw, int t, cf* acc) {
    constexpr int NB = E / R;       // butterflies per thread
    constexpr int M = R * S;        // size of the sub-DFT this pass splits
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int b = t + i * T;
        const int npr = b & (S - 1);
        const int base = (b / S) * M + npr;
        cf a[R];
#pragma unroll
        for (int n = 0; n < R; ++n) a[n] = buf[psg_pad(base + n * S)];
        dftR<R>(a);
        if constexpr (LAST) {
#pragma unroll
            for (int k = 0; k < R; ++k) acc[i * R + k] = fma2(a[k], a[k], acc[i * R + k]);
        } else {
            const float2* twp = tw + npr * (N / M);
#pragma unroll
            for (int k = 1; k < R; ++k) a[k] = cmul(a[k], __ldg(twp + (size_t)npr * (N / M) * (k - 1)));
#pragma unroll
            for (int k = 0; k < R; ++k) buf[psg_pad(base + k * S)] = a[k];
        }
    }
}

template <int LOGN, int E, int R0, int R1, int R2, int R3, int MINB>
__global__ void __launch_bounds__((1 << LOGN) / E, MINB) sti_fused_kernel(const StiArgs a) {
    constexpr int N = 1 << LOGN, T = N / E;
    using PL = Plan<N, R0, R1, R2, R3>;
    constexpr int P = PL::P;
    constexpr int NPAD = psg_pad(N) + 1;
    constexpr int NB0 = E / R0;
    extern __shared__ __align__(16) float2 smem[];

    const int t = threadIdx.x;
    const int item = blockIdx.x;
    const int split = item % a.nsplit;
    const int cs = item / a.nsplit;
    const int col = cs % a.ncol;
    const int sub = cs / a.ncol;
    const int k0 = split * a.chunk;
    const int k1 = min(a.nfr, k0 + a.chunk);

    const float2* src = a.iq + a.col_off[col] + (long long)sub * a.sub_stride + (long long)k0 * a.hop_elems;

    // loop-invariant tables in registers: window of this thread's E samples, pass-0 twiddles
    float w[E];
    cf tw0[NB0][R0 - 1];
#pragma unroll
    for (int i = 0; i < NB0; ++i) {
        const int b = t + i * T;
#pragma unroll
        for (int n = 0; n < R0; ++n) w[i * R0 + n] = __ldg(a.win + b + n * PL::S0);
        if constexpr (P > 1) {
#pragma unroll
            for (int k = 1; k < R0; ++k) tw0[i][k - 1] = __ldg(a.tw + b * k);
        }
    }
    cf acc[E];
#pragma unroll
    for (int i = 0; i < E; ++i) acc[i] = make_float2(0.f, 0.f);

    for (int k = k0; k < k1; ++k, src += a.hop_elems) {
        float2* buf = smem + ((k - k0) & 1) * NPAD;
        cf x[E];
        // ---- pass 0: global -> registers, window, DFT over the most significant digit ----
#pragma unroll
        for (int i = 0; i < NB0; ++i) {
            const int b = t + i * T;
#pragma unroll
            for (int n = 0; n < R0; ++n)
                x[i * R0 + n] = ldg_stream(src + (long long)(b + n * PL::S0) * a.sample_stride);
        }
#pragma unroll
        for (int i = 0; i < E; ++i) x[i] = cscale(x[i], w[i]);
#pragma unroll
        for (int i = 0; i < NB0; ++i) {
            dftR<R0>(&x[i * R0]);
            if constexpr (P == 1) {
#pragma unroll
                for (int kk = 0; kk < R0; ++kk) acc[i * R0 + kk] = fma2(x[i * R0 + kk], x[i * R0 + kk], acc[i * R0 + kk]);
            } else {
                const int b = t + i * T;
#pragma unroll
                for (int kk = 1; kk < R0; ++kk) x[i * R0 + kk] = cmul(x[i * R0 + kk], tw0[i][kk - 1]);
#pragma unroll
                for (int kk = 0; kk < R0; ++kk) buf[psg_pad(b + kk * PL::S0)] = x[i * R0 + kk];
            }
        }
        if constexpr (P >= 2) {
            __syncthreads();
            smem_pass<N, E, T, R1, PL::S1, P == 2>(buf, a.tw, t, acc);
        }
        if constexpr (P >= 3) {
            __syncthreads();
            smem_pass<N, E, T, R2, PL::S2, P == 3>(buf, a.tw, t, acc);
        }
        if constexpr (P >= 4) {
            __syncthreads();
            smem_pass<N, E, T, R3, PL::S3, P == 4>(buf, a.tw, t, acc);
        }
    }

    // ---- epilogue: digit-reversed register sums -> fftshifted, coalesced stores ----
    __syncthreads();
    float* sout = reinterpret_cast<float*>(smem);
    constexpr int RL = PL::RL;
    constexpr int NBL = E / RL;
#pragma unroll
    for (int i = 0; i < NBL; ++i) {
        const int b = t + i * T;
        const int klow = PL::low_freq(b);
#pragma unroll
        for (int j = 0; j < RL; ++j) {
            const int freq = klow + (N / RL) * j;
            const int idx = (freq + N / 2) & (N - 1);
            sout[idx ^ ((idx >> 5) & 31)] = acc[i * RL + j].x + acc[i * RL + j].y;
        }
    }
    __syncthreads();
    if (a.nsplit > 1) {
        float* dst = a.partial + ((size_t)cs * a.nsplit + split) * N;
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const int idx = t + i * T;
            dst[idx] = sout[idx ^ ((idx >> 5) & 31)];
        }
    } else {
        const size_t o = (size_t)cs * N;
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const int idx = t + i * T;
            const float p = sout[idx ^ ((idx >> 5) & 31)] * a.scale;
            if (a.out_lin) a.out_lin[o + idx] = p;
            if (a.out_db) a.out_db[o + idx] = power_to_db(p, a.eps);
        }
    }
}

// Sum the partial columns of split work items (fixed order -> deterministic), scale, store.
__global__ void sti_finalize_kernel(const float* __restrict__ partial, int nsplit, int n, size_t ncols_total,
                                    float scale, float eps, float* out_lin, float* out_db) {
    const size_t total = ncols_total * (size_t)n;
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const size_t c = g / n;
        const int i = (int)(g - c * n);
        const float* p = partial + c * nsplit * (size_t)n + i;
        double s = 0.0;
        for (int k = 0; k < nsplit; ++k) s += (double)p[(size_t)k * n];
        const float v = (float)s * scale;
        if (out_lin) out_lin[g] = v;
        if (out_db) out_db[g] = power_to_db(v, eps);
    }
}

// ---- generic fallback: any power-of-two N that fits shared memory, radix-2, simple ------------
// Used for N < 256, for cross-checking the tuned kernels, and never for speed.
__global__ void __launch_bounds__(256) sti_generic_kernel(const StiArgs a, int logn) {
    const int N = 1 << logn;
    extern __shared__ __align__(16) float2 smem[];
    float2* buf = smem;
    float* accs = reinterpret_cast<float*>(smem + N);
    const int t = threadIdx.x, nt = blockDim.x;
    const int item = blockIdx.x;
    const int split = item % a.nsplit;
    const int cs = item / a.nsplit;
    const int col = cs % a.ncol;
    const int sub = cs / a.ncol;
    const int k0 = split * a.chunk;
    const int k1 = min(a.nfr, k0 + a.chunk);
    const float2* src = a.iq + a.col_off[col] + (long long)sub * a.sub_stride + (long long)k0 * a.hop_elems;
    for (int i = t; i < N; i += nt) accs[i] = 0.f;
    for (int k = k0; k < k1; ++k, src += a.hop_elems) {
        __syncthreads();
        for (int i = t; i < N; i += nt) buf[i] = cscale(ldg_stream(src + (long long)i * a.sample_stride), a.win[i]);
        for (int s = N >> 1; s >= 1; s >>= 1) {
            __syncthreads();
            const int step = (N >> 1) / s;  // twiddle stride: W_{2s}^j = W_N^{j*N/(2s)}
            for (int i = t; i < (N >> 1); i += nt) {
                const int j = i & (s - 1);
                const int base = ((i / s) * 2 * s) + j;
                const cf u = buf[base], v = buf[base + s];
                buf[base] = cadd(u, v);
                cf d = csub(u, v);
                buf[base + s] = (s > 1) ? cmul(d, a.tw[j * step]) : d;
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
