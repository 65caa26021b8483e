// Micro-benchmark: cost of warp-level gathers (ld.global.nc) and scatter-adds (red.global.add.f32)
// as a function of how many 128-byte lines / 32-byte sectors the 32 lanes touch.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/mb_l1 scratch/mb_l1.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// op: 0 = LDG, 1 = RED, 2 = LDG + RED (different arrays, same offsets), 3 = RED with lane pairs on the SAME address
// nl: lines per warp instruction; wstride: word stride between lanes of the same line (1: same sector(s), 8: one sector each)
template <int OP>
__global__ void __launch_bounds__(1024) mb(float *a, float *b, uint32_t nlines_mask, int nl, int wstride, int iters, float *sink) {
    const int lane = threadIdx.x & 31;
    const uint32_t wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int per = 32 / nl;
    int k = lane / per, w = (lane % per) * wstride;
    if (OP == 3) w = ((lane % per) >> 1) * wstride;
    float acc = 0.f;
    for (int it = 0; it < iters; it += 4) {
        uint32_t off[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t line = hash32((wg * 65599u + (uint32_t)(it + u)) * 64u + (uint32_t)k) & nlines_mask;
            off[u] = line * 32u + (uint32_t)w;
        }
        if (OP == 0 || OP == 2) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                float v;
                asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(a + off[u]));
                acc += v;
            }
        }
        if (OP == 1 || OP == 2 || OP == 3) {
#pragma unroll
            for (int u = 0; u < 4; u++)
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(b + off[u]), "f"(1.0f + acc) : "memory");
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

int g_blocks = 148 * 16, g_threads = 128, g_smem = 0;
template <int OP>
void run(const char *name, float *a, float *b, float *sink, size_t nfloats, int nl, int wstride) {
    const int iters = 2048, blocks = g_blocks, threads = g_threads;
    cudaFuncSetAttribute(mb<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem);
    const uint32_t mask = (uint32_t)(nfloats / 32 - 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mb<OP><<<blocks, threads, g_smem>>>(a, b, mask, nl, wstride, 256, sink);
    cudaEventRecord(e0);
    mb<OP><<<blocks, threads, g_smem>>>(a, b, mask, nl, wstride, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double winstr = (double)blocks * (threads / 32) * iters;
    const double cyc = ms * 1e-3 * 1.965e9 * (blocks < 148 ? blocks : 148) / winstr;
    printf("%-10s nl=%2d wstride=%d  %8.3f ms  %7.2f SM-cycles / warp-instr   %6.1f G lane-ops/s\n", name, nl, wstride, ms, cyc,
           winstr * 32 / (ms * 1e-3) / 1e9);
}

int main(int argc, char **argv) {
    if (argc > 1) g_blocks = atoi(argv[1]);
    if (argc > 2) g_threads = atoi(argv[2]);
    if (argc > 3) g_smem = atoi(argv[3]);
    printf("blocks = %d\n", g_blocks);
    for (int pass = 0; pass < 1; pass++) {
        const size_t nfloats = pass == 0 ? (size_t)16 << 20 : (size_t)1 << 18;   // 64 MiB (L2) / 1 MiB
        float *a, *b, *sink;
        cudaMalloc(&a, nfloats * 4); cudaMalloc(&b, nfloats * 4); cudaMalloc(&sink, 4);
        cudaMemset(a, 0, nfloats * 4); cudaMemset(b, 0, nfloats * 4);
        printf("---- array %zu MiB each\n", nfloats * 4 >> 20);
        const int nls[] = {1, 2, 4, 8, 16, 32};
        for (int nl : nls) {
            run<0>("LDG", a, b, sink, nfloats, nl, 1);
            if (32 / nl <= 4) run<0>("LDG", a, b, sink, nfloats, nl, 8);
        }
        for (int nl : nls) {
            run<1>("RED", a, b, sink, nfloats, nl, 1);
            if (32 / nl <= 4) run<1>("RED", a, b, sink, nfloats, nl, 8);
        }
        for (int nl : nls) run<3>("RED-dup2", a, b, sink, nfloats, nl, 1);
        for (int nl : nls) run<2>("LDG+RED", a, b, sink, nfloats, nl, 1);
        cudaFree(a); cudaFree(b); cudaFree(sink);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
