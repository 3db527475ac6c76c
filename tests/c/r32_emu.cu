// CPU replay of the radix-32 whole-frame kernels (pyspectrogram_b200/csrc/sti_r32.cuh) -- test infrastructure.
//
// Compiled by nvcc as a host-only program (no kernel launch, no GPU): it includes the kernel's own math header
// (r32_math.cuh: butterflies, twiddle recurrences, address functions, accumulator-to-bin map; PSG_HD functions are
// bit-identical on the host) and walks the three passes of every geometry thread by thread, exactly as the
// kernel does, on one frame of noise.  Checks
//   1. |X|^2 of every bin against a float64 FFT of the windowed frame (relative to the column peak);
//   2. every address function against the generic swizzle r32_swz of the element it names;
//   3. that every warp-wide shared-memory access of the passes is bank-conflict free.
// Run by tests/test_fft_plan.py (CPU suite).
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../pyspectrogram_b200/csrc/r32_math.cuh"

typedef std::complex<double> cd;

static void fft_ref(std::vector<cd>& a) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = -2.0 * M_PI / (double)len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const cd w(cos(ang * (double)k), sin(ang * (double)k));
                const cd u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

static int g_fail = 0;
#define CHECK(cond, ...)                       \
    do {                                       \
        if (!(cond)) {                         \
            if (g_fail < 20) { printf("FAIL: " __VA_ARGS__); printf("\n"); } \
            ++g_fail;                          \
        }                                      \
    } while (0)

// bank-conflict check of one warp instruction: addr[lane] byte addresses, width bytes per lane (4, 8 or 16)
static void check_banks(const uint32_t* addr, int width, const char* what) {
    const int group = (width == 16) ? 8 : (width == 8) ? 16 : 32;  // lanes served per wavefront
    for (int g0 = 0; g0 < 32; g0 += group) {
        int owner[32];
        for (int b = 0; b < 32; ++b) owner[b] = -1;
        for (int l = g0; l < g0 + group; ++l)
            for (int k = 0; k < width / 4; ++k) {
                const int bank = (int)((addr[l] / 4 + k) & 31);
                CHECK(owner[bank] < 0 || addr[owner[bank]] == addr[l], "%s: lanes %d and %d share bank %d", what, owner[bank], l, bank);
                owner[bank] = l;
            }
    }
}

template <int LOGN, int T>
static void run(unsigned seed) {
    using G = R32Geo<LOGN, T>;
    constexpr int N = G::N, L = G::L, CL = G::CL, NR = G::NR, R1 = G::R1, S1 = G::S1, NB1 = G::NB1, SWSH = G::SWSH;
    srand(seed);
    std::vector<float2> x(N), tw(N);
    std::vector<float> win(N);
    for (int n = 0; n < N; ++n) {
        x[n] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
        win[n] = (float)((0.3 + 0.7 * sin(M_PI * (n + 0.5) / N)) / N);
        const double ang = -2.0 * M_PI * (double)n / (double)N;
        tw[n] = make_float2((float)cos(ang), (float)sin(ang));
    }
    std::vector<std::vector<unsigned char>> M(CL, std::vector<unsigned char>(T * 32 * 8, 0xff));
    std::vector<std::vector<int>> written(CL, std::vector<int>(T * 32, 0));
    auto ld = [&](int cta, uint32_t off) { float2 v; memcpy(&v, &M[cta][off], 8); return v; };
    auto st = [&](int cta, uint32_t off, float2 v) { memcpy(&M[cta][off], &v, 8); };

    // ---- pass 0 ----
    for (int c = 0; c < CL; ++c) {
        for (int w0 = 0; w0 < T; w0 += 32) {
            uint32_t addr[32][32];  // [k0][lane]
            for (int lane = 0; lane < 32; ++lane) {
                const int t = w0 + lane, np = c * T + t;
                cf xr[32], u[16], v[16];
                float wv[32];
                for (int a = 0; a < 32; ++a) xr[a] = x[np + a * L];
                for (int j = 0; j < 16; ++j) { wv[2 * j] = win[np + j * L]; wv[2 * j + 1] = win[np + (j + 16) * L]; }
                dft32_layer8<0, true>(xr, wv, u, v);
                dft32_layer8<8, true>(xr, wv + 16, u, v);
                dft32_finish(xr, u, v);
                cf pw[5];
                for (int q = 0; q < 5; ++q) pw[q] = tw[((unsigned)np << q) & (N - 1)];
                twiddle_dfs32(xr, pw);
                const uint32_t col = r32_p0_col<G>(np);
                for (int k0 = 0; k0 < 32; ++k0) {
                    const int s = k0 / NR, r = k0 % NR;
                    const uint32_t off = col + (uint32_t)r * (L * 8);
                    CHECK(off == r32_swz<SWSH>((uint32_t)(r * L + np)), "pass 0 address N=%d np=%d k0=%d", N, np, k0);
                    st(s, off, xr[k0]);
                    written[s][off / 8]++;
                    addr[k0][lane] = off;
                }
            }
            for (int k0 = 0; k0 < 32; ++k0) check_banks(addr[k0], 8, "pass 0 store");
        }
    }
    for (int c = 0; c < CL; ++c)
        for (int i = 0; i < T * 32; ++i) CHECK(written[c][i] == 1, "pass 0: slot %d of CTA %d written %d times", i, c, written[c][i]);

    // ---- pass 1 ----
    for (int c = 0; c < CL; ++c) {
        for (int w0 = 0; w0 < T; w0 += 32) {
            uint32_t addr[NB1][32][32];
            for (int lane = 0; lane < 32; ++lane) {
                const int t = w0 + lane;
                for (int i = 0; i < NB1; ++i) {
                    const int id = t + i * T, r1 = id / S1, c1 = id & (S1 - 1);
                    const uint32_t base1 = r32_p1_base<G>(t, i);
                    cf xr[32];
                    for (int b = 0; b < R1; ++b) {
                        const uint32_t off = base1 + r32_p1_off<G>(t, b);
                        CHECK(off == r32_swz<SWSH>((uint32_t)(r1 * L + b * S1 + c1)), "pass 1 address N=%d t=%d b=%d", N, t, b);
                        xr[b] = ld(c, off);
                        addr[i][b][lane] = off;
                    }
                    cf pw[5];
                    for (int q = 0; q < 5; ++q) pw[q] = tw[(((unsigned)c1 << q) * 32u) & (N - 1)];
                    if (R1 == 32) {
                        dft32(xr);
                        twiddle_dfs32(xr, pw, [&](int k1, cf v) { st(c, base1 + r32_p1_off<G>(t, k1), v); });
                    } else {
                        dft16(xr);
                        twiddle_dfs16(xr, pw, [&](int k1, cf v) { st(c, base1 + r32_p1_off<G>(t, k1), v); });
                    }
                }
            }
            for (int i = 0; i < NB1; ++i)
                for (int b = 0; b < R1; ++b) check_banks(addr[i][b], 8, "pass 1");
        }
    }

    // ---- pass 2 -> power per bin ----
    std::vector<float> power(N, -1.f);
    for (int c = 0; c < CL; ++c) {
        std::vector<std::vector<cf>> y(T, std::vector<cf>(32));
        std::vector<std::vector<float>> acc(T, std::vector<float>(32, 0.f));
        for (int w0 = 0; w0 < T; w0 += 32) {
            uint32_t addr[32][32];
            for (int lane = 0; lane < 32; ++lane) {
                const int t = w0 + lane;
                if constexpr (S1 == 16) {
                    for (int i = 0; i < 2; ++i) {
                        cf yy[16];
                        int row, k1;
                        r32_p2_block<G>(t, i, row, k1);
                        for (int ch = 0; ch < 8; ++ch) {
                            const uint32_t off = r32_p2_addr<G>(t, i, ch);
                            CHECK(off == r32_swz<SWSH>((uint32_t)(row * L + k1 * 16 + 2 * ch)), "pass 2 address t=%d", t);
                            yy[2 * ch] = ld(c, off);
                            yy[2 * ch + 1] = ld(c, off + 8);
                            addr[i * 8 + ch][lane] = off;
                        }
                        dft16(yy);
                        for (int j = 0; j < 16; ++j) acc[t][16 * i + j] = fmaf(yy[j].x, yy[j].x, yy[j].y * yy[j].y);
                    }
                } else if constexpr (S1 == 32) {
                    cf yy[32];
                    for (int ch = 0; ch < 16; ++ch) {
                        const uint32_t off = r32_p2_addr<G>(t, 0, ch);
                        CHECK(off == r32_swz<SWSH>((uint32_t)((t >> 5) * L + lane * 32 + 2 * ch)), "pass 2 address t=%d", t);
                        yy[2 * ch] = ld(c, off);
                        yy[2 * ch + 1] = ld(c, off + 8);
                        addr[ch][lane] = off;
                    }
                    dft32(yy);
                    for (int j = 0; j < 32; ++j) acc[t][j] = fmaf(yy[j].x, yy[j].x, yy[j].y * yy[j].y);
                } else {
                    const int r2 = t >> 6, k1 = (t & 63) >> 1, e = t & 1;
                    for (int d = 0; d < 32; ++d) {
                        const uint32_t off = r32_p2_addr<G>(t, 0, d);
                        CHECK(off == r32_swz<SWSH>((uint32_t)(r2 * L + 64 * k1 + 2 * d + e)), "pass 2 address t=%d", t);
                        y[t][d] = ld(c, off);
                        addr[d][lane] = off;
                    }
                    dft32(y[t].data());
                }
            }
            const int ninstr = (S1 == 64) ? 32 : 16;
            for (int i = 0; i < ninstr; ++i) check_banks(addr[i], S1 == 64 ? 8 : 16, "pass 2");
        }
        if constexpr (S1 == 64) {
            for (int t = 0; t < T; ++t) {
                const int e = t & 1;
                for (int i = 0; i < 16; ++i) {
                    const cf keep = e ? y[t][16 + i] : y[t][i];
                    const cf recv = e ? y[t ^ 1][16 + i] : y[t ^ 1][i];  // what the partner (e' = 1 - e) sends: its y[e' ? i : 16 + i]
                    cf s0, s1;
                    r32_pair_finish(e, i, keep, recv, s0, s1);
                    acc[t][2 * i] = fmaf(s0.x, s0.x, s0.y * s0.y);
                    acc[t][2 * i + 1] = fmaf(s1.x, s1.x, s1.y * s1.y);
                }
            }
        }
        // the epilogue's staging: conflict check of the scalar stores, then the bin map
        for (int t = 0; t < T; ++t)
            for (int ai = 0; ai < 32; ++ai) {
                int r, m;
                r32_acc_bin<G>(t, ai, r, m);
                const int freq = (c * NR + r) + 32 * m;
                CHECK(freq >= 0 && freq < N && power[freq] < 0.f, "bin map N=%d t=%d ai=%d -> %d", N, t, ai, freq);
                power[freq] = acc[t][ai];
            }
        for (int w0 = 0; w0 < T; w0 += 32)
            for (int ai = 0; ai < 32; ++ai) {
                uint32_t addr[32];
                for (int lane = 0; lane < 32; ++lane) {
                    int r, m;
                    r32_acc_bin<G>(w0 + lane, ai, r, m);
                    addr[lane] = 4u * (uint32_t)(r * G::RS + ((m + N / 64) & (N / 32 - 1)));
                }
                if (S1 != 64 && R1 == 32) check_banks(addr, 4, "epilogue staging store");  // S1 = 64 and R1 = 16: two-way by design (once per work item)
            }
    }

    // ---- reference ----
    std::vector<cd> ref(N);
    for (int n = 0; n < N; ++n) ref[n] = cd((double)x[n].x * win[n], (double)x[n].y * win[n]);
    fft_ref(ref);
    double peak = 0, worst = 0, worst_rel = 0;
    for (int k = 0; k < N; ++k) peak = std::max(peak, std::norm(ref[k]));
    for (int k = 0; k < N; ++k) {
        const double p = std::norm(ref[k]);
        worst = std::max(worst, fabs((double)power[k] - p) / peak);
        if (p > 1e-3 * peak) worst_rel = std::max(worst_rel, fabs((double)power[k] - p) / p);
    }
    printf("N=%d T=%d CL=%d: max |err| / peak = %.3g, max rel err (bins > -30 dB) = %.3g\n", N, T, CL, worst, worst_rel);
    CHECK(worst < 2e-6, "N=%d: error %.3g", N, worst);
    CHECK(worst_rel < 1e-5, "N=%d: relative error %.3g", N, worst_rel);
}

int main() {
    run<13, 256>(1);
    run<14, 512>(2);
    run<14, 256>(3);
    run<15, 512>(4);
    run<15, 256>(5);
    run<16, 512>(6);
    printf(g_fail ? "FAILED (%d)\n" : "OK\n", g_fail);
    return g_fail ? 1 : 0;
}
