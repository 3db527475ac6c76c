// Register-level math and index algebra of the radix-32 whole-frame kernels (sti_r32.cuh).  Everything here
// compiles for the host as well (PSG_HD), bit-identically, so tests/c/r32_emu.cu replays the three passes of
// every geometry on the CPU with the kernel's own butterflies, twiddle recurrences and address functions.
#pragma once
#include <stdint.h>
#include "cplx.cuh"

// ---- 32-point DFT in registers ----------------------------------------------------------------------------
// W_32^j = exp(-2 pi i j / 32)
template <int J>
PSG_HD cf w32_const() {
    constexpr float c[9] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f, 0.70710678118654752440f,
                            0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f, 0.f};
    constexpr int j = J & 31;
    // cos(pi j / 16) by symmetry from the first octant-and-a-half table
    constexpr int jc = (j <= 8) ? j : (j <= 16) ? 16 - j : (j <= 24) ? j - 16 : 32 - j;
    constexpr float cs = (j <= 8 || j >= 24) ? c[jc] : -c[jc];
    constexpr int js = (j + 24) & 31;  // sin(x) = cos(x - pi/2)
    constexpr int jsc = (js <= 8) ? js : (js <= 16) ? 16 - js : (js <= 24) ? js - 16 : 32 - js;
    constexpr float sn = (js <= 8 || js >= 24) ? c[jsc] : -c[jsc];
    return make_float2(cs, -sn);
}
template <int J>
PSG_HD cf mul_w32(cf d) {
    if constexpr ((J & 31) == 0) return d;
    else if constexpr ((J & 31) == 8) return mul_nj(d);
    else return cmul(d, w32_const<J>());
}
// radix-2 layer (j, j + 16) with W_32^j on the differences, then two 16-point DFTs: even / odd outputs.
// WIN: the inputs are multiplied by w[2 j] (element j) and w[2 j + 1] (element j + 16) inside that layer
// (the TMEM column order of the window), 3 instead of 4 instructions per pair.
template <int J, bool WIN>
PSG_HD void dft32_pair(const cf* a, const float* w, cf* u, cf* v) {
    cf d;
    if constexpr (WIN) {
        wpair(a[J], w[2 * (J & 7)], a[J + 16], w[2 * (J & 7) + 1], u[J], d);
    } else {
        u[J] = cadd(a[J], a[J + 16]);
        d = csub(a[J], a[J + 16]);
    }
    v[J] = mul_w32<J>(d);
}
template <int J0, bool WIN>
PSG_HD void dft32_layer8(const cf* a, const float* w, cf* u, cf* v) {  // pairs J0 .. J0 + 7; w = their 16 window values
    dft32_pair<J0 + 0, WIN>(a, w, u, v);
    dft32_pair<J0 + 1, WIN>(a, w, u, v);
    dft32_pair<J0 + 2, WIN>(a, w, u, v);
    dft32_pair<J0 + 3, WIN>(a, w, u, v);
    dft32_pair<J0 + 4, WIN>(a, w, u, v);
    dft32_pair<J0 + 5, WIN>(a, w, u, v);
    dft32_pair<J0 + 6, WIN>(a, w, u, v);
    dft32_pair<J0 + 7, WIN>(a, w, u, v);
}
PSG_HD void dft32_finish(cf* a, cf* u, cf* v) {
    dft16(u);
    dft16(v);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        a[2 * m] = u[m];
        a[2 * m + 1] = v[m];
    }
}
PSG_HD void dft32(cf* a) {
    cf u[16], v[16];
    dft32_layer8<0, false>(a, nullptr, u, v);
    dft32_layer8<8, false>(a, nullptr, u, v);
    dft32_finish(a, u, v);
}

// x[k] *= W^k, k = 1..31, from pw[q] = W^(2^q), q < 5, depth first (see tw_visit); every finished output is
// handed to emit(k, x[k]) at once, so the stores of a pass interleave with its twiddle arithmetic instead of
// forming a pure load/store phase behind it (the LSU phases are what the FMA pipe idles on, tools/r32_trace.py)
template <int K, int QMIN, class F>
PSG_HD void tw32_visit(cf* x, const cf* pw, const cf tk, F& emit) {
    x[K] = cmul(x[K], tk);
    emit(K, x[K]);
    if constexpr (QMIN <= 0 && K + 1 < 32) tw32_visit<K + 1, 1>(x, pw, cmul(tk, pw[0]), emit);
    if constexpr (QMIN <= 1 && K + 2 < 32) tw32_visit<K + 2, 2>(x, pw, cmul(tk, pw[1]), emit);
    if constexpr (QMIN <= 2 && K + 4 < 32) tw32_visit<K + 4, 3>(x, pw, cmul(tk, pw[2]), emit);
    if constexpr (QMIN <= 3 && K + 8 < 32) tw32_visit<K + 8, 4>(x, pw, cmul(tk, pw[3]), emit);
    if constexpr (QMIN <= 4 && K + 16 < 32) tw32_visit<K + 16, 5>(x, pw, cmul(tk, pw[4]), emit);
}
template <class F>
PSG_HD void twiddle_dfs32(cf* x, const cf* pw, F&& emit) {
    emit(0, x[0]);
    tw32_visit<1, 1>(x, pw, pw[0], emit);
    tw32_visit<2, 2>(x, pw, pw[1], emit);
    tw32_visit<4, 3>(x, pw, pw[2], emit);
    tw32_visit<8, 4>(x, pw, pw[3], emit);
    tw32_visit<16, 5>(x, pw, pw[4], emit);
}
// the same for a 16-point butterfly (pw[q], q < 4)
template <int K, int QMIN, class F>
PSG_HD void tw16_visit(cf* x, const cf* pw, const cf tk, F& emit) {
    x[K] = cmul(x[K], tk);
    emit(K, x[K]);
    if constexpr (QMIN <= 0 && K + 1 < 16) tw16_visit<K + 1, 1>(x, pw, cmul(tk, pw[0]), emit);
    if constexpr (QMIN <= 1 && K + 2 < 16) tw16_visit<K + 2, 2>(x, pw, cmul(tk, pw[1]), emit);
    if constexpr (QMIN <= 2 && K + 4 < 16) tw16_visit<K + 4, 3>(x, pw, cmul(tk, pw[2]), emit);
    if constexpr (QMIN <= 3 && K + 8 < 16) tw16_visit<K + 8, 4>(x, pw, cmul(tk, pw[3]), emit);
}
template <class F>
PSG_HD void twiddle_dfs16(cf* x, const cf* pw, F&& emit) {
    emit(0, x[0]);
    tw16_visit<1, 1>(x, pw, pw[0], emit);
    tw16_visit<2, 2>(x, pw, pw[1], emit);
    tw16_visit<4, 3>(x, pw, pw[2], emit);
    tw16_visit<8, 4>(x, pw, pw[3], emit);
}
struct R32NoEmit {
    PSG_HD void operator()(int, cf) const {}
};
PSG_HD void twiddle_dfs32(cf* x, const cf* pw) { twiddle_dfs32(x, pw, R32NoEmit{}); }

// W_64^i, i < 16 (the radix-2 butterfly of the L = 2048 rows); i is a compile-time constant after unrolling
PSG_HD cf w64_table(int i) {
    constexpr float c[17] = {1.f, 0.99518472667219688624f, 0.98078528040323044913f, 0.95694033573220886494f,
                             0.92387953251128675613f, 0.88192126434835502971f, 0.83146961230254523708f, 0.77301045336273696081f,
                             0.70710678118654752440f, 0.63439328416364549822f, 0.55557023301960222474f, 0.47139673682599764856f,
                             0.38268343236508977173f, 0.29028467725446236764f, 0.19509032201612826785f, 0.09801714032956060199f, 0.f};
    return make_float2(c[i], -c[16 - i]);
}

// byte offset inside M of element pos (complex index inside this CTA's rows)
template <int SWSH>
PSG_HD constexpr uint32_t r32_swz(uint32_t pos) {
    const uint32_t line = pos >> 4;
    return line * 128u + ((((pos >> 1) & 7u) ^ ((line >> SWSH) & 7u)) << 4) + ((pos & 1u) << 3);
}

// ---- geometry (independent of the sample type) ------------------------------------------------------------------
// N = 2^LOGN = 32 L; a CTA of T threads owns T columns n' of pass 0 and NR = 32 T / L rows of length L afterwards;
// CL = L / T CTAs (one cluster) share a frame.  Rows: L = R1 S1, pass 1 radix R1 at stride S1, pass 2 the S1-point
// DFTs on consecutive positions (S1 = 16: radix 16; 32: radix 32; 64: radix 32 at stride 2 + a lane-pair butterfly).
//   8192  = 32 x 256 (T = 256, two CTAs per SM)      rows 16 x 16
//   16384 = 32 x 512 (T = 512; or T = 256 on pairs)   rows 32 x 16
//   32768 = 32 x 1024 (T = 512 on pairs)              rows 32 x 32
//   65536 = 32 x 2048 (T = 512 on clusters of four)   rows 32 x 32 x 2
template <int LOGN, int T_>
struct R32Geo {
    static constexpr int N = 1 << LOGN, T = T_, L = N / 32, CL = L / T, NR = 32 / CL;
    static constexpr int R1 = (L >= 512) ? 32 : 16, S1 = L / R1, NB1 = 32 / R1;
    static constexpr int SWSH = (S1 == 16) ? 0 : (S1 == 32) ? 1 : 2;  // line-index bits XORed into the chunk index
    static constexpr int RS = L + CL;  // row stride of the epilogue's staging, in floats
    static_assert(L % T == 0 && (CL == 1 || CL == 2 || CL == 4 || CL == 8) && NR * CL == 32, "cluster shape");
    static_assert(S1 == 16 || S1 == 32 || S1 == 64, "row plan");
    static_assert(T % 32 == 0 && NR * S1 == T * NB1, "NB1 pass-1 butterflies per thread");
};

// Byte offsets inside M.  Each is r32_swz(pos) of the element it names, written so that everything but the
// thread-dependent part folds into an immediate once the loops are unrolled (checked against r32_swz on the CPU).
// pass 0: output k0 of column n' (n' = T c + t) -> row r = k0 % NR of CTA k0 / NR, element r L + n'
template <class G>
PSG_HD uint32_t r32_p0_col(int np) { return r32_swz<G::SWSH>((uint32_t)np); }  // + r * L * 8
// pass 1: butterfly i of thread t is (row r1, c1) = divmod(t + i T, S1), element r1 L + b S1 + c1; swizzle bits b & 7
template <class G>
PSG_HD uint32_t r32_p1_base(int t, int i) {
    const int id = t + i * G::T, r1 = id / G::S1, c1 = id & (G::S1 - 1);
    return (uint32_t)r1 * (G::L * 8) + (uint32_t)(c1 >> 4) * 128u + (uint32_t)(c1 & 1) * 8u;
}
template <class G>
PSG_HD uint32_t r32_p1_off(int t, int b) {
    const uint32_t cc = (uint32_t)(((t & (G::S1 - 1)) & 15) >> 1);  // c1 does not depend on i: T is a multiple of S1
    return (uint32_t)b * (G::S1 * 8) + ((cc ^ (uint32_t)(b & 7)) << 4);
}
// pass 2
//   S1 = 16: a warp's pass-1 rows hold 64 blocks of 16 consecutive elements (one line each); lane takes blocks
//            jj = lane + 32 i (i < 2): row 2 w + (jj / R1 >> 1) T / 16 + (jj / R1 & 1), block k1 = jj % R1;
//            j = 16-byte chunk (elements 2 j, 2 j + 1)
//   S1 = 32: row w, block k1 = lane: 32 consecutive elements = two lines; j = chunk < 16
//   S1 = 64: row t / 64, k1 = (t % 64) / 2, e = t & 1: elements 64 k1 + 2 j + e, j < 32 (64-bit accesses)
template <class G>
PSG_HD void r32_p2_block(int t, int i, int& row, int& k1) {  // S1 = 16
    const int lane = t & 31, w = t >> 5, jj = lane + 32 * i, sel = jj / G::R1;
    row = 2 * w + (sel >> 1) * (G::T / 16) + (sel & 1);
    k1 = jj % G::R1;
}
template <class G>
PSG_HD uint32_t r32_p2_addr(int t, int i, int j) {
    constexpr int L = G::L;
    const int lane = t & 31, w = t >> 5;
    if constexpr (G::S1 == 16) {
        int row, k1;
        r32_p2_block<G>(t, i, row, k1);
        return (uint32_t)row * (L * 8) + (uint32_t)k1 * 128u + (((uint32_t)j ^ (uint32_t)(k1 & 7)) << 4);
    } else if constexpr (G::S1 == 32) {
        return (uint32_t)w * (L * 8) + (uint32_t)lane * 256u + (uint32_t)(j >> 3) * 128u + (((uint32_t)(j & 7) ^ (uint32_t)(lane & 7)) << 4);
    } else {
        const int r2 = t >> 6, k1 = (t & 63) >> 1, e = t & 1;
        return (uint32_t)r2 * (L * 8) + (uint32_t)k1 * 512u + (uint32_t)e * 8u + (uint32_t)(j >> 3) * 128u +
               (((uint32_t)(j & 7) ^ (uint32_t)(k1 & 7)) << 4);
    }
}
// accumulator ai of thread t holds the bin  freq = (c NR + r) + 32 m  of the frame
template <class G>
PSG_HD void r32_acc_bin(int t, int ai, int& r, int& m) {
    const int lane = t & 31, w = t >> 5;
    if constexpr (G::S1 == 16) {
        int k1;
        r32_p2_block<G>(t, ai >> 4, r, k1);
        m = k1 + G::R1 * (ai & 15);
    } else if constexpr (G::S1 == 32) {
        r = w;
        m = lane + 32 * ai;
    } else {
        r = t >> 6;
        m = ((t & 63) >> 1) + 32 * (((t & 1) ? 16 : 0) + (ai >> 1) + 32 * (ai & 1));
    }
}
// the radix-2 butterfly between the lanes of a pair (S1 = 64): from this lane's 32 outputs y of the stride-2
// DFT and the partner's (recv = the partner's y[e ? i : 16 + i]), the two outputs lane e finishes for i
PSG_HD void r32_pair_finish(int e, int i, cf keep, cf recv, cf& s0, cf& s1) {
    const cf y0 = e ? recv : keep;
    cf z = cmul(e ? keep : recv, w64_table(i));
    if (e) z = mul_nj(z);
    s0 = cadd(y0, z);
    s1 = csub(y0, z);
}
