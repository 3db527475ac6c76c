// Pipe-throughput microbenchmarks for sm_100a (tuning aid, not product code).
// Each kernel runs ITER iterations of 16 independent instructions per thread; we report
// warp-instructions per clock per SM from cudaEvent time and the measured SM clock (clock64).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../pyspectrogram_b200/csrc/cplx.cuh"

#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, float s, long long* clk) {
    cf a[16];
    float f[16];
    for (int i = 0; i < 16; ++i) { a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f); f[i] = a[i].x; }
    cf w = make_float2(s, 1.0f - s);
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) f[i] = fmaf(f[i], s, 0.25f * s + f[(i + 1) & 15] * 0.f);          // scalar FFMA x1 (+1 FFMA)
            if (MODE == 1) a[i] = fma2(a[i], w, w);                                          // FFMA2
            if (MODE == 2) a[i] = cadd(a[i], w);                                             // FADD2
            if (MODE == 3) { a[i] = fma2(a[i], w, w); f[i] = fmaf(f[i], s, s); }             // 1 FFMA2 + 1 FFMA
            if (MODE == 4) { a[i] = fma2(a[i], w, w); f[i] = fmaf(f[i], s, s); f[i] = fmaf(f[i], s, w.y); }  // 1 + 2
            if (MODE == 5) f[i] = fmaf(f[i], s, s);                                          // scalar FFMA
            if (MODE == 6) { a[i] = cadd(a[i], w); f[i] = f[i] + s; }                         // FADD2 + FADD
        }
    }
    long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < 16; ++i) acc += a[i].x + a[i].y + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(acc, 0);
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

// shared-memory throughput: MODE 0 LDS.64, 1 LDS.128, 2 STS.64, 3 LDS.64+STS.64
template <int MODE>
__global__ void __launch_bounds__(256) ks(float2* out, long long* clk) {
    __shared__ __align__(16) float2 sm[256 * 17];
    for (int i = threadIdx.x; i < 256 * 17; i += 256) sm[i] = make_float2(i, 0);
    __syncthreads();
    float2 acc = make_float2(0, 0);
    float4 acc4 = make_float4(0, 0, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int idx = threadIdx.x + i * 256 + ((it & 1) ? 1 : 0);
            const unsigned sa = (unsigned)__cvta_generic_to_shared(&sm[idx]);
            if (MODE == 0) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(sa)); acc.x += v.x; acc.y += v.y; }
            if (MODE == 1) { float4 v; const unsigned sb = (unsigned)__cvta_generic_to_shared(&sm[(2 * threadIdx.x + i * 512) % (256 * 16)]);
                             asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sb)); acc4.x += v.x; acc4.y += v.w; }
            if (MODE == 2) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(sa), "f"(acc.x), "f"(acc.y)); }
            if (MODE == 4) { float v; const unsigned sc = (unsigned)__cvta_generic_to_shared(reinterpret_cast<float*>(sm) + (threadIdx.x + i * 256 + (it & 1)) % (256 * 32));
                             asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sc)); acc.x += v; }
            if (MODE == 5) { const unsigned sb = (unsigned)__cvta_generic_to_shared(&sm[(2 * threadIdx.x + i * 512) % (256 * 16)]);
                             asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(sb), "f"(acc.x), "f"(acc.y), "f"(acc.x), "f"(acc.y)); }
            if (MODE == 6) { float4 v; const unsigned sb = (unsigned)__cvta_generic_to_shared(&sm[(2 * threadIdx.x + i * 512) % (256 * 16)]);
                             asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sb));
                             asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(sb), "f"(v.y), "f"(v.x), "f"(v.w), "f"(v.z)); }
            if (MODE == 3) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(sa));
                             asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(sa), "f"(v.y), "f"(v.x)); }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(acc.x + acc4.x, acc.y + acc4.y);
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <typename F>
void run(const char* name, F launch, int instr_per_iter_elem, int blocks_per_sm) {
    float2* out; long long* clk;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2)); cudaMalloc(&clk, 8);
    launch(out, clk, blocks_per_sm);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); launch(out, clk, blocks_per_sm); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    double winstr = (double)ITER * 16 * instr_per_iter_elem * 8 /*warps*/ * blocks_per_sm;
    printf("%-28s blocks/SM=%d  cycles=%lld  warp-instr/clk/SM=%.3f  (%.3f ms)\n", name, blocks_per_sm, c, winstr / c, ms);
    cudaFree(out); cudaFree(clk);
}

int main() {
    for (int bps : {2}) {
        run("FFMA (dep chain +1)", [](float2* o, long long* c, int b) { k<0><<<148 * b, 256>>>(o, 0.5f, c); }, 2, bps);
        run("FFMA scalar", [](float2* o, long long* c, int b) { k<5><<<148 * b, 256>>>(o, 0.5f, c); }, 1, bps);
        run("FFMA2", [](float2* o, long long* c, int b) { k<1><<<148 * b, 256>>>(o, 0.5f, c); }, 1, bps);
        run("FADD2", [](float2* o, long long* c, int b) { k<2><<<148 * b, 256>>>(o, 0.5f, c); }, 1, bps);
        run("FFMA2 + FFMA", [](float2* o, long long* c, int b) { k<3><<<148 * b, 256>>>(o, 0.5f, c); }, 2, bps);
        run("FFMA2 + 2 FFMA", [](float2* o, long long* c, int b) { k<4><<<148 * b, 256>>>(o, 0.5f, c); }, 3, bps);
        run("FADD2 + FADD", [](float2* o, long long* c, int b) { k<6><<<148 * b, 256>>>(o, 0.5f, c); }, 2, bps);
        run("LDS.64", [](float2* o, long long* c, int b) { ks<0><<<148 * b, 256>>>(o, c); }, 1, bps);
        run("LDS.128", [](float2* o, long long* c, int b) { ks<1><<<148 * b, 256>>>(o, c); }, 1, bps);
        run("STS.64", [](float2* o, long long* c, int b) { ks<2><<<148 * b, 256>>>(o, c); }, 1, bps);
        run("LDS.64+STS.64", [](float2* o, long long* c, int b) { ks<3><<<148 * b, 256>>>(o, c); }, 2, bps);
        run("LDS.32", [](float2* o, long long* c, int b) { ks<4><<<148 * b, 256>>>(o, c); }, 1, bps);
        run("STS.128", [](float2* o, long long* c, int b) { ks<5><<<148 * b, 256>>>(o, c); }, 1, bps);
        run("LDS.128+STS.128", [](float2* o, long long* c, int b) { ks<6><<<148 * b, 256>>>(o, c); }, 2, bps);
    }
    return 0;
}
