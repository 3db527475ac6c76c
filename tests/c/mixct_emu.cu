// CPU replay of the compile-time mixed-radix plans (sti_mixct.cuh): the prime-factor butterflies and the whole
// pass structure (index algebra, twiddles, digit reversal) against a float64 DFT.  Host code only: the
// butterflies are the PSG_HD functions the kernel calls.   nvcc -O2 -std=c++17 -I pyspectrogram_b200/csrc -o mixct_emu tests/c/mixct_emu.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <complex>
#include <vector>

#include "sti_mixct.cuh"

typedef std::complex<double> cd;
static double frand() { return rand() / (double)RAND_MAX - 0.5; }

template <int R>
static double check_butterfly() {
    cf v[R];
    cd x[R];
    for (int i = 0; i < R; ++i) {
        x[i] = cd(frand(), frand());
        v[i] = make_float2((float)x[i].real(), (float)x[i].imag());
        x[i] = cd(v[i].x, v[i].y);
    }
    mx_dft<R>(v);
    double err = 0, nrm = 0;
    for (int k = 0; k < R; ++k) {
        cd s = 0;
        for (int n = 0; n < R; ++n) s += x[n] * std::polar(1.0, -2 * M_PI * n * k / R);
        err = fmax(err, std::abs(s - cd(v[k].x, v[k].y)));
        nrm = fmax(nrm, std::abs(s));
    }
    return err / nrm;
}

template <class PL, int PIDX>
static void host_pass(std::vector<cf>& buf, const std::vector<cf>& tw) {
    constexpr int N = PL::N, RR = PL::r(PIDX), Sp = PL::s(PIDX);
    for (int bf = 0; bf < N / RR; ++bf) {
        const int blk = bf / Sp, npr = bf - blk * Sp, base = blk * RR * Sp + npr;
        cf v[RR];
        cf* const pb = buf.data() + PL::pad(base);
        for (int n = 0; n < RR; ++n) {
            if (PL::pad(base) + PL::off(PIDX, n) != PL::pad(base + n * Sp)) { printf("padding not linear: N=%d pass %d\n", N, PIDX); exit(3); }
            v[n] = pb[PL::off(PIDX, n)];
        }
        mx_dft<RR>(v);
        if (PIDX < PL::P - 1) {
            constexpr int ts = N / (RR * Sp), NPW = mx_npow(RR);
            if (PL::TWREG) {
                for (int k = 1; k < RR; ++k) v[k] = cmul(v[k], tw[npr * k * ts]);
            } else {
                cf pw[NPW];
                for (int q = 0; q < NPW; ++q) pw[q] = tw[((npr * ts) << q) % N];
                mx_twiddle_from_powers<RR>(v, pw);
            }
        }
        for (int k = 0; k < RR; ++k) pb[PL::off(PIDX, k)] = v[k];
    }
}

template <class PL>
static double check_plan() {
    constexpr int N = PL::N;
    std::vector<cf> tw(N), buf(PL::BUF + 64);
    std::vector<cd> x(N);
    for (int m = 0; m < N; ++m) tw[m] = make_float2((float)cos(2 * M_PI * m / N), (float)-sin(2 * M_PI * m / N));
    for (int n = 0; n < N; ++n) {
        const cf s = make_float2((float)frand(), (float)frand());
        x[n] = cd(s.x, s.y);
        buf[PL::pad(n)] = s;
    }
    host_pass<PL, 0>(buf, tw);
    host_pass<PL, 1>(buf, tw);
    if constexpr (PL::P > 2) host_pass<PL, 2>(buf, tw);
    if constexpr (PL::P > 3) host_pass<PL, 3>(buf, tw);
    // float64 DFT at a few hundred bins
    double err = 0, nrm = 0;
    std::vector<int> at(N, -1);
    for (int pos = 0; pos < N; ++pos) {
        const int f = PL::freq(pos);
        if (f < 0 || f >= N || at[f] != -1) return 1e9;  // not a permutation
        at[f] = pos;
    }
    for (int k = 0; k < N; k += (N > 2000 ? 37 : 7)) {
        cd s = 0;
        for (int n = 0; n < N; ++n) s += x[n] * std::polar(1.0, -2 * M_PI * (double)((long long)n * k % N) / N);
        const cf g = buf[PL::pad(at[k])];
        err = fmax(err, std::abs(s - cd(g.x, g.y)));
        nrm = fmax(nrm, std::abs(s));
    }
    return err / nrm;
}

#define MIXCT_ALT(ID, N, R0, R1, R2, R3, T, PQ, PA, TW, FD, MINB, PQ2, PA2) MIXCT_PLAN(N, R0, R1, R2, R3, T, PQ, PA, TW, FD, MINB, PQ2, PA2)
#define MIXCT_PLAN(N, R0, R1, R2, R3, T, PQ, PA, TW, FD, MINB, PQ2, PA2) \
    { const double e = check_plan<MixPlan<N, R0, R1, R2, R3, T, PQ, PA, TW, PQ2, PA2>>(); printf("plan %6d %2dx%2dx%2dx%2d  err/peak %.2e\n", N, R0, R1, R2, R3, e); bad |= !(e < 3e-6); }

int main() {
    int bad = 0;
#define BF(R) { const double e = check_butterfly<R>(); printf("dft%-2d err/peak %.2e\n", R, e); bad |= !(e < 1e-6); }
    BF(2) BF(3) BF(4) BF(5) BF(6) BF(8) BF(10) BF(12) BF(15) BF(16) BF(20)
#include "mixct_plans.inc"
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad;
}
