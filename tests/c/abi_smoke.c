/* Pure-C caller of libpsgb200.so: only <stdint.h>/<math.h> and include/psg_b200.h -- no CUDA
 * headers, no torch, no Python.  What a C/C++ host (or any FFI) does to use the path:
 *   plan -> psg_sti_host (host buffers in, host buffers out) -> compare with a naive float64 DFT.
 * Build: gcc -O2 -I include tests/c/abi_smoke.c -o abi_smoke -L pyspectrogram_b200 -lpsgb200 -lm
 * (tests/test_gpu_parity.py::test_c_abi_from_plain_c does this and runs it on the GPU box).
 * Follows drfProc.py:386-401: Kaiser(1.7) periodic window, |FFT(w x)|^2 / sum(w)^2, fftshift,
 * median over time; and drfProc.py:308-310 for the dB image. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "psg_b200.h"

#define NFFT 64
#define NCOL 5
#define NINT 3

static double i0(double x) {
    double q = 0.25 * x * x, term = 1.0, sum = 1.0;
    for (int k = 1; k < 200; ++k) { term *= q / ((double)k * k); sum += term; }
    return sum;
}

static int cmp_float(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

int main(void) {
    const double pi = 3.14159265358979323846;
    const int n = NFFT * NINT * NCOL + 7;
    float* iq = (float*)malloc(sizeof(float) * 2 * n);
    uint32_t s = 12345u;
    for (int i = 0; i < 2 * n; ++i) {
        s = s * 1664525u + 1013904223u;
        iq[i] = ((float)(s >> 8) / 16777216.0f - 0.5f) * 0.02f;
    }
    for (int i = 0; i < n; ++i) { /* a tone at bin 5 */
        iq[2 * i] += 0.1f * (float)cos(2 * pi * 5.0 * i / NFFT);
        iq[2 * i + 1] += 0.1f * (float)sin(2 * pi * 5.0 * i / NFFT);
    }
    int64_t starts[NCOL];
    for (int c = 0; c < NCOL; ++c) starts[c] = (int64_t)c * NFFT * NINT + (c & 1);

    psg_plan* plan = NULL;
    if (psg_version() != PSG_ABI_VERSION) { printf("ABI version mismatch\n"); return 2; }
    int rc = psg_plan_create(&plan, NFFT, PSG_WINDOW_KAISER, 1.7, 0);
    if (rc) { printf("psg_plan_create: %d %s\n", rc, psg_last_error()); return 3; }
    static float lin[NCOL * NFFT], db[NCOL * NFFT], med[NFFT], med_db[NFFT];
    rc = psg_sti_host(plan, iq, n, 1, 0, 1, starts, NCOL, NINT, NFFT, 1.0f, 1e-15f, lin, db, med, med_db);
    if (rc) { printf("psg_sti_host: %d %s\n", rc, psg_last_error()); return 4; }

    /* float64 reference */
    double w[NFFT], wsum = 0.0;
    for (int k = 0; k < NFFT; ++k) {
        double r = (k - 0.5 * NFFT) / (0.5 * NFFT);
        w[k] = i0(1.7 * sqrt(1.0 - r * r)) / i0(1.7);
        wsum += w[k];
    }
    double worst = 0.0, worst_db = 0.0;
    static float ref_img[NCOL][NFFT];
    for (int c = 0; c < NCOL; ++c) {
        double col[NFFT], peak = 0.0;
        for (int k = 0; k < NFFT; ++k) col[k] = 0.0;
        for (int f = 0; f < NINT; ++f) {
            const float* x = iq + 2 * (starts[c] + (int64_t)f * NFFT);
            for (int k = 0; k < NFFT; ++k) {
                double re = 0.0, im = 0.0;
                for (int m = 0; m < NFFT; ++m) {
                    double a = -2 * pi * (double)((k * m) % NFFT) / NFFT, cr = cos(a), ci = sin(a);
                    double xr = x[2 * m] * w[m], xi = x[2 * m + 1] * w[m];
                    re += xr * cr - xi * ci;
                    im += xr * ci + xi * cr;
                }
                col[(k + NFFT / 2) % NFFT] += (re * re + im * im) / (wsum * wsum) / NINT;
            }
        }
        for (int k = 0; k < NFFT; ++k) if (col[k] > peak) peak = col[k];
        for (int k = 0; k < NFFT; ++k) {
            double e = fabs(lin[c * NFFT + k] - col[k]) / peak;
            if (e > worst) worst = e;
            if (col[k] >= peak * 1e-6) {
                double d = fabs(db[c * NFFT + k] - 10.0 * log10((double)(float)col[k] + 1e-15));
                if (d > worst_db) worst_db = d;
            }
            ref_img[c][k] = lin[c * NFFT + k];
        }
    }
    /* median over time of the library's own image must be exact */
    int med_bad = 0;
    for (int k = 0; k < NFFT; ++k) {
        float v[NCOL];
        for (int c = 0; c < NCOL; ++c) v[c] = ref_img[c][k];
        qsort(v, NCOL, sizeof(float), cmp_float);
        if (v[NCOL / 2] != med[k]) ++med_bad;
    }
    printf("c-abi smoke: max err/peak %.3e, max dB err %.3e, median mismatches %d, kernel %s, launches %lld\n", worst,
           worst_db, med_bad, psg_plan_variant(plan), (long long)psg_launch_count());
    psg_plan_destroy(plan);
    free(iq);
    return (worst <= 1e-5 && worst_db <= 1e-3 && med_bad == 0) ? 0 : 1;
}
