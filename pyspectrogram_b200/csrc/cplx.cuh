// Packed complex fp32 arithmetic for sm_100a and register-resident DFT butterflies.
//
// Blackwell adds packed fp32x2 ALU ops (PTX add/sub/mul/fma .f32x2 -> SASS FADD2 / FMUL2 /
// FFMA2).  ptxas folds half-swaps (".LO_HI"), per-half negation and scalar broadcast (".F32")
// into the operand modifiers, so with (re, im) kept in one 64-bit register pair:
//     complex add/sub         = 1 instruction
//     multiply by +-j         = free (operand swizzle of the consuming add)
//     real scalar * complex   = 1 instruction
//     complex * complex       = 2 instructions (FMUL2 + FFMA2)
//     |x|^2 accumulate        = 1 instruction (acc.re += re^2, acc.im += im^2)
// That halves the issue slots of the FFT against scalar FADD/FFMA code.
#pragma once
#include <cuda_runtime.h>

#define PSG_DEV __device__ __forceinline__
// arithmetic that also compiles for the host (bit-identical: fused multiply-adds, round to nearest), so the
// butterflies and the index algebra of a kernel can be replayed on the CPU (tests/c/r32_emu.cu)
#define PSG_HD __host__ __device__ __forceinline__

typedef float2 cf;

PSG_HD cf cadd(cf a, cf b) {
#ifdef __CUDA_ARCH__
    cf r;
    asm("{ .reg .b64 ra, rb, rd;\n\t mov.b64 ra, {%2, %3};\n\t mov.b64 rb, {%4, %5};\n\t"
        " add.rn.f32x2 rd, ra, rb;\n\t mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
PSG_HD cf csub(cf a, cf b) {
#ifdef __CUDA_ARCH__
    cf r;
    asm("{ .reg .b64 ra, rb, rd;\n\t mov.b64 ra, {%2, %3};\n\t mov.b64 rb, {%4, %5};\n\t"
        " sub.rn.f32x2 rd, ra, rb;\n\t mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
PSG_HD cf mul2(cf a, cf b) {  // elementwise (a.x*b.x, a.y*b.y)
#ifdef __CUDA_ARCH__
    cf r;
    asm("{ .reg .b64 ra, rb, rd;\n\t mov.b64 ra, {%2, %3};\n\t mov.b64 rb, {%4, %5};\n\t"
        " mul.rn.f32x2 rd, ra, rb;\n\t mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
PSG_HD cf fma2(cf a, cf b, cf c) {  // elementwise a*b+c
#ifdef __CUDA_ARCH__
    cf r;
    asm("{ .reg .b64 ra, rb, rc, rd;\n\t mov.b64 ra, {%2, %3};\n\t mov.b64 rb, {%4, %5};\n\t"
        " mov.b64 rc, {%6, %7};\n\t fma.rn.f32x2 rd, ra, rb, rc;\n\t mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
PSG_HD cf cscale(cf a, float s) { return mul2(a, make_float2(s, s)); }
// a * w  (complex)
PSG_HD cf cmul(cf a, cf w) {
    cf t = mul2(a, make_float2(w.x, w.x));
    return fma2(make_float2(a.y, a.x), make_float2(-w.y, w.y), t);
}
// a * conj(w)
PSG_HD cf cmulc(cf a, cf w) {
    cf t = mul2(a, make_float2(w.x, w.x));
    return fma2(make_float2(a.y, a.x), make_float2(w.y, -w.y), t);
}
PSG_HD cf mul_nj(cf a) { return make_float2(a.y, -a.x); }  // a * (-j)
PSG_HD cf mul_pj(cf a) { return make_float2(-a.y, a.x); }  // a * (+j)

// ---- forward DFT butterflies, in place, natural-order outputs (exp(-2*pi*j*n*k/R)) -------------

PSG_HD void dft2(cf& a0, cf& a1) {
    cf s = cadd(a0, a1);
    a1 = csub(a0, a1);
    a0 = s;
}

PSG_HD void dft4(cf& a0, cf& a1, cf& a2, cf& a3) {
    cf t0 = cadd(a0, a2), t1 = csub(a0, a2);
    cf t2 = cadd(a1, a3), d = csub(a1, a3);
    cf t3 = mul_nj(d);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

// c + a*w (complex): 2 instructions
PSG_HD cf cfma(cf a, cf w, cf c) {
    cf t = fma2(a, make_float2(w.x, w.x), c);
    return fma2(make_float2(a.y, a.x), make_float2(-w.y, w.y), t);
}
PSG_HD cf twice_minus(cf m, cf s) {  // 2*m - s: the difference of a butterfly from its sum, 1 instruction
    return fma2(m, make_float2(2.0f, 2.0f), make_float2(-s.x, -s.y));
}

// 4-point DFT of (a0, W1*v1, W2*v2, W3*v3) with the constant multiplies folded into the first
// layer: 12 instructions instead of 3 complex multiplies + 8 adds = 14.
PSG_HD void dft4_tw(cf& a0, cf& v1, cf& v2, cf& v3, cf W1, cf W2, cf W3) {
    const cf t0 = cfma(v2, W2, a0), t1 = twice_minus(a0, t0);
    const cf m1 = cmul(v1, W1);
    const cf t2 = cfma(v3, W3, m1), d = twice_minus(m1, t2);
    const cf t3 = mul_nj(d);
    a0 = cadd(t0, t2);
    v2 = csub(t0, t2);
    v1 = cadd(t1, t3);
    v3 = csub(t1, t3);
}
// same with W2 = -j (free): 11 instead of 12
PSG_HD void dft4_tw_nj(cf& a0, cf& v1, cf& v2, cf& v3, cf W1, cf W3) {
    const cf r2 = mul_nj(v2);
    const cf t0 = cadd(a0, r2), t1 = csub(a0, r2);
    const cf m1 = cmul(v1, W1);
    const cf t2 = cfma(v3, W3, m1), d = twice_minus(m1, t2);
    const cf t3 = mul_nj(d);
    a0 = cadd(t0, t2);
    v2 = csub(t0, t2);
    v1 = cadd(t1, t3);
    v3 = csub(t1, t3);
}

#define PSG_SQRT1_2 0.70710678118654752440f
#define PSG_C1_16 0.92387953251128675613f  // cos(pi/8)
#define PSG_S1_16 0.38268343236508977173f  // sin(pi/8)

PSG_HD void dft8(cf* a) {
    // radix-2 over (i, i+4), twiddle W8^i on the differences, then two 4-point DFTs
    cf s0 = cadd(a[0], a[4]), d0 = csub(a[0], a[4]);
    cf s1 = cadd(a[1], a[5]), d1 = csub(a[1], a[5]);
    cf s2 = cadd(a[2], a[6]), d2 = csub(a[2], a[6]);
    cf s3 = cadd(a[3], a[7]), d3 = csub(a[3], a[7]);
    dft4(s0, s1, s2, s3);  // X[0], X[2], X[4], X[6]
    // X[1], X[3], X[5], X[7]: W8^1, W8^2 = -j, W8^3 on d1, d2, d3 folded into the 4-point DFT
    dft4_tw_nj(d0, d1, d2, d3, make_float2(PSG_SQRT1_2, -PSG_SQRT1_2), make_float2(-PSG_SQRT1_2, -PSG_SQRT1_2));
    a[0] = s0; a[2] = s1; a[4] = s2; a[6] = s3;
    a[1] = d0; a[3] = d1; a[5] = d2; a[7] = d3;
}

PSG_HD void dft16(cf* a) {
    // DIF 4x4: X[c + 4d] = sum_i W4^{i d} ( W16^{i c} sum_m a[4m + i] W4^{m c} )
    cf u[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        u[i][0] = a[i]; u[i][1] = a[i + 4]; u[i][2] = a[i + 8]; u[i][3] = a[i + 12];
        dft4(u[i][0], u[i][1], u[i][2], u[i][3]);  // index c
    }
    // internal twiddles W16^{i*c}
    // internal twiddles W16^{i*c} folded into the second layer of 4-point DFTs (index d)
    dft4(u[0][0], u[1][0], u[2][0], u[3][0]);
    dft4_tw(u[0][1], u[1][1], u[2][1], u[3][1], make_float2(PSG_C1_16, -PSG_S1_16),      // W16^1
            make_float2(PSG_SQRT1_2, -PSG_SQRT1_2),                                        // W16^2
            make_float2(PSG_S1_16, -PSG_C1_16));                                           // W16^3
    dft4_tw_nj(u[0][2], u[1][2], u[2][2], u[3][2], make_float2(PSG_SQRT1_2, -PSG_SQRT1_2),  // W16^2, (W16^4 = -j)
               make_float2(-PSG_SQRT1_2, -PSG_SQRT1_2));                                   // W16^6
    dft4_tw(u[0][3], u[1][3], u[2][3], u[3][3], make_float2(PSG_S1_16, -PSG_C1_16),      // W16^3
            make_float2(-PSG_SQRT1_2, -PSG_SQRT1_2),                                       // W16^6
            make_float2(-PSG_C1_16, PSG_S1_16));                                           // W16^9
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        a[c] = u[0][c]; a[c + 4] = u[1][c]; a[c + 8] = u[2][c]; a[c + 12] = u[3][c];
    }
}

// Windowed variants for the first pass: a[n] = x[n] * w[n] folded into the first radix-2 layer.
// (x0 w0 + x2 w2, x0 w0 - x2 w2) = (fma(x2, w2, m), 2 m - (..)) with m = x0 w0: 3 instructions per
// pair instead of 4 (two multiplies, add, subtract).
PSG_HD void wpair(cf x0, float w0, cf x2, float w2, cf& s, cf& d) {
    const cf m = mul2(x0, make_float2(w0, w0));
    s = fma2(x2, make_float2(w2, w2), m);
    d = fma2(m, make_float2(2.0f, 2.0f), make_float2(-s.x, -s.y));
}

PSG_HD void dft4w(cf& a0, cf& a1, cf& a2, cf& a3, float w0, float w1, float w2, float w3) {
    cf t0, t1, t2, d;
    wpair(a0, w0, a2, w2, t0, t1);
    wpair(a1, w1, a3, w3, t2, d);
    cf t3 = mul_nj(d);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

PSG_HD void dft16w(cf* a, const float* w) {
    cf u[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        u[i][0] = a[i]; u[i][1] = a[i + 4]; u[i][2] = a[i + 8]; u[i][3] = a[i + 12];
        dft4w(u[i][0], u[i][1], u[i][2], u[i][3], w[i], w[i + 4], w[i + 8], w[i + 12]);
    }
    // internal twiddles W16^{i*c} folded into the second layer of 4-point DFTs (index d)
    dft4(u[0][0], u[1][0], u[2][0], u[3][0]);
    dft4_tw(u[0][1], u[1][1], u[2][1], u[3][1], make_float2(PSG_C1_16, -PSG_S1_16),      // W16^1
            make_float2(PSG_SQRT1_2, -PSG_SQRT1_2),                                        // W16^2
            make_float2(PSG_S1_16, -PSG_C1_16));                                           // W16^3
    dft4_tw_nj(u[0][2], u[1][2], u[2][2], u[3][2], make_float2(PSG_SQRT1_2, -PSG_SQRT1_2),  // W16^2, (W16^4 = -j)
               make_float2(-PSG_SQRT1_2, -PSG_SQRT1_2));                                   // W16^6
    dft4_tw(u[0][3], u[1][3], u[2][3], u[3][3], make_float2(PSG_S1_16, -PSG_C1_16),      // W16^3
            make_float2(-PSG_SQRT1_2, -PSG_SQRT1_2),                                       // W16^6
            make_float2(-PSG_C1_16, PSG_S1_16));                                           // W16^9
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        a[c] = u[0][c]; a[c + 4] = u[1][c]; a[c + 8] = u[2][c]; a[c + 12] = u[3][c];
    }
}

// window multiply + R-point DFT; the fold is implemented for R = 4 and 16 (first passes of the
// 1024- and 4096-point plans), other radices multiply first
template <int R>
PSG_HD void dftRw(cf* a, const float* w) {
    if constexpr (R == 16) dft16w(a, w);
    else if constexpr (R == 4) dft4w(a[0], a[1], a[2], a[3], w[0], w[1], w[2], w[3]);
    else {
#pragma unroll
        for (int n = 0; n < R; ++n) a[n] = mul2(a[n], make_float2(w[n], w[n]));
        if constexpr (R == 2) dft2(a[0], a[1]);
        else if constexpr (R == 8) dft8(a);
    }
}

template <int R>
PSG_HD void dftR(cf* a) {
    if constexpr (R == 2) dft2(a[0], a[1]);
    else if constexpr (R == 4) dft4(a[0], a[1], a[2], a[3]);
    else if constexpr (R == 8) dft8(a);
    else if constexpr (R == 16) dft16(a);
}
