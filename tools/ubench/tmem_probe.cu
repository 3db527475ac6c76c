// TMEM as a per-thread scratch file (tuning aid, not product code).
// Checks the assumptions the radix-32 whole-frame kernels rely on (sti_r32.cuh):
//   1. tcgen05.alloc of all 512 columns from a 512-thread CTA that never issues an MMA;
//   2. warp w reads / writes lanes 32*(w % 4) .. +31 with the .32x32b shapes: thread i of the warp owns
//      lane 32*(w%4)+i, columns col .. col+x-1 -> 128 private 32-bit words per thread at 16 warps;
//   3. tcgen05.st -> wait::st -> tcgen05.ld -> wait::ld round trip returns what was stored;
//   4. throughput of tcgen05.ld / tcgen05.st (x8, x32) with 16 resident warps, and of LDS.64 next to it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define DEV __device__ __forceinline__

DEV void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
DEV void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
DEV void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
        "[%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
DEV void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::
            "r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
DEV void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
DEV void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE 0: correctness round trip; 1: ld x32 loop; 2: ld x8 loop; 3: st x32 loop; 4: LDS.64 loop; 5: LDS.64 + ld x8 interleaved
template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(uint32_t* out, long long* clk, int iters) {
    __shared__ uint32_t tbase_s;
    __shared__ __align__(16) float2 sm[512 * 9];
    const int t = threadIdx.x, w = t >> 5;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tbase_s)),
                     "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = t; i < 512 * 9; i += 512) sm[i] = make_float2((float)i, 1.f);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tbase_s;
    // warp w: lanes 32*(w&3).., columns 128*(w>>2) .. +127
    const uint32_t mine = tbase + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(128 * (w >> 2));
    uint32_t bad = 0;
    long long t0 = 0, t1 = 0;
    if (MODE == 0) {
        uint32_t r[32];
        for (int c = 0; c < 128; c += 32) {
            for (int i = 0; i < 32; ++i) r[i] = (uint32_t)(t * 1000 + c + i);
            tmem_st32(mine + c, r);
        }
        tmem_wait_st();
        // read back with x8 at every 8-column offset
        for (int c = 0; c < 128; c += 8) {
            uint32_t q[8];
            tmem_ld8(mine + c, q);
            tmem_wait_ld();
            for (int i = 0; i < 8; ++i) bad += (q[i] != (uint32_t)(t * 1000 + c + i));
        }
        // overwrite a slice with x8, read with x32
        uint32_t q8[8];
        for (int i = 0; i < 8; ++i) q8[i] = 7u * t + i;
        tmem_st8(mine + 40, q8);
        tmem_wait_st();
        tmem_ld32(mine + 32, r);
        tmem_wait_ld();
        for (int i = 0; i < 32; ++i) {
            const uint32_t want = (i >= 8 && i < 16) ? 7u * t + (i - 8) : (uint32_t)(t * 1000 + 32 + i);
            bad += (r[i] != want);
        }
    } else {
        uint32_t r[32];
        for (int i = 0; i < 32; ++i) r[i] = t + i;
        tmem_st32(mine, r);
        tmem_st32(mine + 32, r);
        tmem_wait_st();
        uint32_t acc = 0;
        float2 facc = make_float2(0.f, 0.f);
        __syncthreads();
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (MODE == 1) {
                tmem_ld32(mine + 32 * (it & 1), r);
                tmem_wait_ld();
                acc += r[it & 31];
            }
            if (MODE == 2) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t q[8];
                    tmem_ld8(mine + 8 * c, q);
                    tmem_wait_ld();
                    acc += q[it & 7];
                }
            }
            if (MODE == 3) {
                r[0] = acc + it;
                tmem_st32(mine + 32 * (it & 1), r);
                tmem_wait_st();
            }
            if (MODE == 4 || MODE == 5) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 v = sm[t + i * 512 + (it & 1)];
                    facc.x += v.x;
                    facc.y += v.y;
                }
                if (MODE == 5) {
                    uint32_t q[8];
                    tmem_ld8(mine + 8 * (it & 3), q);
                    tmem_wait_ld();
                    acc += q[it & 7];
                }
            }
        }
        t1 = clock64();
        bad = acc + (uint32_t)facc.x + (uint32_t)facc.y;
    }
    out[blockIdx.x * 512 + t] = bad;
    if (t == 0 && blockIdx.x == 0) *clk = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512));
}

template <int MODE>
static void run(const char* what, int iters, double bytes_per_iter_per_thread) {
    uint32_t* out;
    long long* clk;
    cudaMalloc(&out, 148 * 512 * 4);
    cudaMalloc(&clk, 8);
    probe<MODE><<<148, 512>>>(out, clk, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%-28s CUDA error: %s\n", what, cudaGetErrorString(e));
        exit(1);
    }
    long long c = 0;
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    if (MODE == 0) {
        static uint32_t h[148 * 512];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long bad = 0;
        for (int i = 0; i < 148 * 512; ++i) bad += h[i];
        printf("%-28s mismatches: %llu\n", what, bad);
    } else {
        printf("%-28s %lld clk / %d iters = %.1f clk/iter; %.1f B/clk/SM\n", what, c, iters, (double)c / iters,
               bytes_per_iter_per_thread * 512.0 * iters / (double)c);
    }
    cudaFree(out);
    cudaFree(clk);
}

int main() {
    run<0>("round trip", 1, 0);
    run<1>("tcgen05.ld x32", 4096, 128);
    run<2>("tcgen05.ld 4 * x8", 4096, 128);
    run<3>("tcgen05.st x32", 4096, 128);
    run<4>("8 LDS.64", 4096, 64);
    run<5>("8 LDS.64 + tcgen05.ld x8", 4096, 64);
    return 0;
}
