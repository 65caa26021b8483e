// tc05_probe.cu -- stand-alone probe of the tcgen05 pieces the tensor-core MV-CNN needs (run under gpurun):
// kind::tf32 MMA with SW128 K-major operands written by hand, descriptors whose start is shifted by whole rows
// (the kx tap of a 3x3 convolution) with / without base_offset, N = 32 and 64, TMEM alloc + tcgen05.ld.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scratch/tc05_probe scratch/tc05_probe.cu && ./scratch/tc05_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define ROWS_A 136
#define KDIM 32

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t base_offset) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);            // start address
    d |= (uint64_t)1 << 16;                           // LBO = 1 (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // SBO = 1024 B between 8-row groups
    d |= (uint64_t)1 << 46;                           // version = 1 (Blackwell)
    d |= (uint64_t)(base_offset & 7) << 49;
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                 // C format F32
    d |= 2u << 7;                 // A format TF32
    d |= 2u << 10;                // B format TF32
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

__global__ void __launch_bounds__(128) probe(const float *A, const float *B, float *D, int N, int kx, int use_base_offset) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    unsigned char *sA = smem;                       // ROWS_A x 128 B
    unsigned char *sB = smem + 18 * 1024;           // up to 64 x 128 B (1024-aligned)
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < ROWS_A * 8; i += 128) {   // 16-byte chunks, address-based 128B swizzle
        const int p = i >> 3, c = i & 7;
        *reinterpret_cast<float4 *>(sA + p * 128 + ((c ^ (p & 7)) << 4)) = *reinterpret_cast<const float4 *>(A + p * KDIM + c * 4);
    }
    for (int i = tid; i < N * 8; i += 128) {
        const int p = i >> 3, c = i & 7;
        *reinterpret_cast<float4 *>(sB + p * 128 + ((c ^ (p & 7)) << 4)) = *reinterpret_cast<const float4 *>(B + p * KDIM + c * 4);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, N);
        for (int k = 0; k < 4; k++) {
            const uint32_t a_addr = smem_u32(sA) + kx * 128 + k * 32;
            const uint32_t b_addr = smem_u32(sB) + k * 32;
            const uint64_t da = make_desc(a_addr, use_base_offset ? ((a_addr >> 7) & 7) : 0);
            const uint64_t db = make_desc(b_addr, 0);
            const uint32_t acc = k > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {   // everybody waits for the MMAs
        uint32_t done;
        do {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        } while (!done);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                       "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                       "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; j++) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

int main() {
    std::vector<float> A(ROWS_A * KDIM), B(64 * KDIM);
    srand(1);
    for (auto &x : A) x = (float)((rand() % 17) - 8) / 8.0f;      // exactly representable in tf32: products exact in fp32
    for (auto &x : B) x = (float)((rand() % 17) - 8) / 4.0f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 128 * 64 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    int fails = 0;
    for (int N : {32, 64})
        for (int kx = 0; kx < 3; kx++)
            for (int ubo = 0; ubo < 2; ubo++) {
                cudaMemset(dD, 0xff, 128 * 64 * 4);
                probe<<<1, 128, 32 * 1024>>>(dA, dB, dD, N, kx, ubo);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("N=%d kx=%d base_offset=%d: CUDA error %s\n", N, kx, ubo, cudaGetErrorString(e)); return 1; }
                std::vector<float> D(128 * N);
                cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
                double worst = 0;
                for (int m = 0; m < 128; m++)
                    for (int n = 0; n < N; n++) {
                        double ref = 0;
                        for (int k = 0; k < KDIM; k++) ref += (double)A[(m + kx) * KDIM + k] * (double)B[n * KDIM + k];
                        worst = fmax(worst, fabs(ref - (double)D[m * N + n]));
                    }
                printf("N=%d kx=%d base_offset=%s: max |D - ref| = %g %s\n", N, kx, ubo ? "(addr>>7)&7" : "0", worst, worst == 0 ? "OK" : "MISMATCH");
                if (worst != 0 && (kx == 0)) fails++;
            }
    printf("probe done, kx=0 failures: %d\n", fails);
    return 0;
}
