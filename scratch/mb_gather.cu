// Ceiling of the similarity kernel's access pattern: LDG.128 where 8 lanes cover one random 128-byte vector (4 vectors per
// warp instruction), 16 loads in flight per warp, 32 warps per SM, table of `mb` MiB (L1-, L2- or DRAM-resident).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a scratch/mb_gather.cu -o gpurun_out/mb_gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 8) gather(const float4 *tab, uint32_t mask, int iters, float *out, int per_cta_window) {
    const int lane = threadIdx.x & 31;
    uint32_t s = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 2654435761u + (lane >> 3) * 40503u + 12345u;
    const uint32_t base = per_cta_window ? (blockIdx.x % 148) * (mask + 1) : 0;
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        float4 f[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            s = s * 1664525u + 1013904223u;
            const uint32_t vec = base + ((s >> 8) & mask);
            f[j] = __ldg(tab + (size_t)vec * 8 + (lane & 7));
        }
#pragma unroll
        for (int j = 0; j < 16; j++) acc += f[j].x + f[j].y + f[j].z + f[j].w;
    }
    if (acc == 123.456f) out[0] = acc;
}
int main() {
    const size_t bytes = (size_t)1 << 31;
    float4 *tab; float *out;
    cudaMalloc(&tab, bytes); cudaMemset(tab, 0, bytes); cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    struct { const char *name; uint32_t vecs; int window; } cases[] = {
        {"L1-resident (32 KB window per SM)", 256, 1}, {"L1-resident (96 KB window per SM)", 768 - 1, 1}, {"L2-resident (32 MiB)", 1u << 18, 0},
        {"L2-resident (64 MiB)", 1u << 19, 0}, {"mostly DRAM (2 GiB)", 1u << 24, 0}, {"316 MB like the C3 feature maps", (1u << 21) + (1u << 19), 0}};
    for (auto &c : cases) {
        uint32_t mask = 1; while (mask * 2 <= c.vecs) mask *= 2; mask -= 1;
        const int iters = 64, blocks = 148 * 8 * 4;
        gather<<<blocks, 128>>>(tab, mask, 8, out, c.window);
        cudaEventRecord(e0);
        gather<<<blocks, 128>>>(tab, mask, iters, out, c.window);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double ldg = (double)blocks * 4 * iters * 16, bytes_moved = ldg * 512;
        printf("%-40s table %7.1f MiB  %.3f ms  %.2f TB/s  %.1f B/clk/SM at %d MHz nominal  %.2f ns per LDG.128 per SM\n", c.name, (mask + 1) * 128.0 / 1048576, ms,
               bytes_moved / ms * 1e-9, bytes_moved / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000, ms * 1e6 / (ldg / 148));
    }
    return 0;
}
