// rn_cnn_tc.cuh -- the 32 -> 32 channel layers of the MV-CNN (models.py:90-111) on the
// 5th-generation tensor cores: an implicit-GEMM 3x3 convolution with tcgen05.mma kind::tf32, accumulators in TMEM.
//
// Precision: the BP marginals downstream are gated at 1e-5, which single-pass TF32 products (10-bit mantissa) do not
// meet.  Every operand travels as an exact pair x = hi + lo (hi = x with its low 13 mantissa bits cleared, i.e.
// exactly representable in TF32; lo = x - hi, exact in float32) and a product is formed as
//     hi_a hi_b + hi_a lo_b + lo_a hi_b                       ("3 x TF32": error ~ 2^-21 of a product, float32 sums)
// Activations are kept as hi / lo arrays between the layers (the epilogue of a layer writes both), weights are split
// on the host.  The dropped lo_a lo_b term is below 2^-22.
//
// GEMM view of one output row segment of 128 pixels:  D[128 px][32 cout] += A_tap[128 px][32 cin] . B_tap[32 cout][32 cin]^T
// over the 9 taps.  A_tap is not copied: an input row segment (130 pixels x 32 channels, one 128-byte shared-memory
// row per pixel, 128B-swizzled) is loaded ONCE and the tap (ky, kx) is the same buffer read through a shared-memory
// descriptor whose start address is shifted by kx rows (scratch/tc05_probe.cu: the swizzle is a function of the
// address, so the shifted view is consistent; base_offset stays 0).  Vertically a CTA walks down a strip, so every
// input row is loaded once and used by three output rows (ring of 4 row stages).
// MMAs per output row segment: per tap and 8-channel slice (kind::tf32 has K = 8) one N = 64 instruction
//     A_hi . [B_hi ; B_lo]^T  ->  TMEM columns 0..31 (hi hi) and 32..63 (hi lo)
// and one N = 32 instruction A_lo . B_hi^T accumulated into columns 0..31: 72 instructions, two thirds of the
// operand traffic of three separate products.  The epilogue adds the two column halves.
//
// Warp roles (192 threads, one CTA per SM, persistent over (image, 128-pixel strip, 32-row chunk) units):
//   warp 0      producer: one lane issues two TMA tensor copies (cp.async.bulk.tensor.2d, 128B swizzle) per input row
//               (hi and lo) into the ring.  (A first version copied with per-lane 16-byte cp.async: its 130 issue
//               slots per row sat on the critical path between the completion of one row's MMAs and the next.)
//   warp 1      one elected lane issues the MMAs; tcgen05.commit releases ring stages / publishes accumulators
//   warps 2-5   epilogue: tcgen05.ld (each warp its 32-lane quarter), folded batch norm (+ ReLU), hi / lo split,
//               128-byte rows to HBM; two accumulator stages in TMEM so the epilogue overlaps the next row's MMAs
// mbarriers: full[4] (producer -> MMA), empty[4] (MMA -> producer, by tcgen05.commit), acc_full[2], acc_empty[2].
#pragma once

#include <cuda.h>      // CUtensorMap (the encode function is fetched through cudaGetDriverEntryPoint, no libcuda link)

#include "rn_common.cuh"

#define RN_TC_PX 128                      // output pixels per MMA (M)
#define RN_TC_ROWS 136                    // shared-memory rows per stage half (130 used: 128 + 2 halo pixels), 1024-byte multiple
#define RN_TC_STAGES 4
#define RN_TC_CHUNK_ROWS 32               // output rows per work unit
#define RN_TC_HALF_BYTES (RN_TC_ROWS * 128)          // 17408 = 17 * 1024
#define RN_TC_STAGE_BYTES (2 * RN_TC_HALF_BYTES)     // hi + lo
#define RN_TC_B_BYTES (9 * 64 * 128)                 // per tap [B_hi ; B_lo]: 64 rows of 128 bytes
#define RN_TC_SMEM_BYTES (RN_TC_STAGES * RN_TC_STAGE_BYTES + RN_TC_B_BYTES + 1024)

struct ConvTcArgs {
    const float *in_hi, *in_lo;   // [N][Hi][Wi][32]
    const float *w_cat;           // [9 taps][64][32]: rows 0..31 = hi part of W[tap][cout][cin], rows 32..63 = lo part
    const float *scale, *shift;   // [32]
    float *out_hi, *out_lo;       // [N][Hi-2][Wi-2][32]; out_lo null: out_hi receives the plain float32 result
    int n, hi, wi, relu;
};

__device__ __forceinline__ uint32_t rn_tc_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rn_tc_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rn_tc_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rn_tc_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rn_tc_smem(bar)) : "memory");
}
__device__ __forceinline__ void rn_tc_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(rn_tc_smem(bar)), "r"(parity) : "memory");
    } while (!done);
}
// K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t rn_tc_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor: C = F32, A = B = TF32, both K-major, M x N
__host__ __device__ constexpr uint32_t rn_tc_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void rn_tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void rn_tc_commit(uint64_t *bar) {   // arrives on bar when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(rn_tc_smem(bar)) : "memory");
}
// one lane of a converged warp (elect.sync: the form ptxas keeps warp-uniform, so that descriptors computed from
// warp-uniform values stay in uniform registers instead of being moved there lane by lane)
__device__ __forceinline__ bool rn_tc_elect() {
    uint32_t pred;
    asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void rn_tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}

// work unit u -> (image, strip x0, first output row, rows)
struct TcUnit {
    int img, x0, oy0, rows;
};
__device__ __forceinline__ bool rn_tc_unit(const ConvTcArgs &a, int64_t u, TcUnit &t) {
    const int ho = a.hi - 2, wo = a.wi - 2;
    const int strips = (wo + RN_TC_PX - 1) / RN_TC_PX, chunks = (ho + RN_TC_CHUNK_ROWS - 1) / RN_TC_CHUNK_ROWS;
    if (u >= (int64_t)a.n * strips * chunks) return false;
    t.img = (int)(u / (strips * chunks));
    const int rem = (int)(u - (int64_t)t.img * strips * chunks);
    t.x0 = (rem / chunks) * RN_TC_PX;
    t.oy0 = (rem % chunks) * RN_TC_CHUNK_ROWS;
    t.rows = min(RN_TC_CHUNK_ROWS, ho - t.oy0);
    return true;
}

__global__ void __launch_bounds__(192, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_hi,
                                                            const __grid_constant__ CUtensorMap tm_lo, ConvTcArgs a) {
    extern __shared__ unsigned char rn_tc_smem_raw[];
    __shared__ __align__(8) uint64_t full[RN_TC_STAGES], empty[RN_TC_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float sScale[32], sShift[32];
    // 1024-byte aligned carve-up: ring stages (hi half, lo half), then the weights
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(rn_tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sB = base + RN_TC_STAGES * RN_TC_STAGE_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // weights: [tap][64 rows][32 k] -> swizzled K-major rows of 128 bytes
    for (int i = tid; i < 9 * 64 * 8; i += blockDim.x) {
        const int row = i >> 3, c = i & 7;                    // row = tap * 64 + n
        const float4 v = __ldg(reinterpret_cast<const float4 *>(a.w_cat) + i);
        *reinterpret_cast<float4 *>(sB + row * 128 + ((c ^ (row & 7)) << 4)) = v;
    }
    if (tid < 32) { sScale[tid] = __ldg(a.scale + tid); sShift[tid] = __ldg(a.shift + tid); }
    if (tid == 0) {
        for (int s = 0; s < RN_TC_STAGES; s++) { rn_tc_mbar_init(&full[s], 1); rn_tc_mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; s++) { rn_tc_mbar_init(&acc_full[s], 1); rn_tc_mbar_init(&acc_empty[s], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // 2 accumulator stages x 64 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rn_tc_smem(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the weights were written through the generic proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        // ================= producer =================
        // One lane, two TMA tensor copies per input row (hi and lo): a box of 130 pixels x 32 channels of the
        // [pixels][32] view of the activations lands as 130 swizzled 128-byte rows (CU_TENSOR_MAP_SWIZZLE_128B writes
        // exactly the layout the MMA descriptors read); completion is counted in bytes on the stage's mbarrier.  Boxes
        // that run past the end of an image row read the next row's pixels (never used), past the end of the tensor zeros.
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_hi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_lo) : "memory");
            int ri = 0;
            TcUnit t;
            for (int64_t u = blockIdx.x; rn_tc_unit(a, u, t); u += gridDim.x) {
                const int pix0 = (t.img * a.hi + t.oy0) * a.wi + t.x0;
                for (int r = 0; r < t.rows + 2; r++, ri++) {
                    const int stage = ri % RN_TC_STAGES;
                    rn_tc_mbar_wait(&empty[stage], ((ri / RN_TC_STAGES) & 1) ^ 1);
                    const uint32_t dst_hi = rn_tc_smem(base + stage * RN_TC_STAGE_BYTES), dst_lo = dst_hi + RN_TC_HALF_BYTES;
                    const uint32_t bar = rn_tc_smem(&full[stage]);
                    const int pix = pix0 + r * a.wi;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * (RN_TC_PX + 2) * 128) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(dst_hi), "l"(&tm_hi), "r"(0), "r"(pix), "r"(bar) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(dst_lo), "l"(&tm_lo), "r"(0), "r"(pix), "r"(bar) : "memory");
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp runs the loop (warp-uniform control flow and addresses: the descriptors stay in uniform
        // registers); only the tcgen05 instructions themselves are issued by one lane.
        constexpr uint32_t idesc64 = rn_tc_idesc(RN_TC_PX, 64), idesc32 = rn_tc_idesc(RN_TC_PX, 32);
        const uint64_t desc_b0 = rn_tc_desc(rn_tc_smem(sB));
        const uint64_t desc_a0 = rn_tc_desc(rn_tc_smem(base));
        int ri_base = 0, tile = 0;
        TcUnit t;
        for (int64_t u = blockIdx.x; rn_tc_unit(a, u, t); u += gridDim.x) {
            for (int j = 0; j < t.rows; j++, tile++) {
                const int as = tile & 1;
                rn_tc_mbar_wait(&acc_empty[as], ((tile >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem + (uint32_t)(as * 64);
                uint32_t accumulate = 0;
                for (int ky = 0; ky < 3; ky++) {
                    const int row = ri_base + j + ky, stage = row % RN_TC_STAGES;
                    rn_tc_mbar_wait(&full[stage], (row / RN_TC_STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // descriptors differ in their 14-bit start-address field only (units of 16 bytes)
                    const uint64_t da_hi = desc_a0 + (uint64_t)((stage * RN_TC_STAGE_BYTES) >> 4);
                    const uint64_t da_lo = da_hi + (uint64_t)(RN_TC_HALF_BYTES >> 4);
                    const uint64_t db_ky = desc_b0 + (uint64_t)((ky * 3 * 64 * 128) >> 4);
                    if (rn_tc_elect()) {
#pragma unroll
                        for (int kx = 0; kx < 3; kx++) {
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const uint64_t sh = (uint64_t)((kx * 128 + k * 32) >> 4);
                                const uint64_t db = db_ky + (uint64_t)((kx * 64 * 128 + k * 32) >> 4);
                                rn_tc_mma(d_tmem, da_hi + sh, db, idesc64, accumulate);   // hi hi | hi lo
                                rn_tc_mma(d_tmem, da_lo + sh, db, idesc32, 1);            // + lo hi into columns 0..31
                                accumulate = 1;
                            }
                        }
                    }
                    __syncwarp();
                }
                if (rn_tc_elect()) {
                    rn_tc_commit(&acc_full[as]);                                  // accumulators of this row are complete
                    rn_tc_commit(&empty[(ri_base + j) % RN_TC_STAGES]);          // input row j is not needed any more
                    if (j == t.rows - 1) {
                        rn_tc_commit(&empty[(ri_base + j + 1) % RN_TC_STAGES]);
                        rn_tc_commit(&empty[(ri_base + j + 2) % RN_TC_STAGES]);
                    }
                }
                __syncwarp();
            }
            ri_base += t.rows + 2;
        }
    } else {
        // ================= epilogue =================
        const int q = warp & 3;                 // this warp's TMEM lane quarter
        const int px = q * 32 + lane;           // pixel of the segment = TMEM lane
        const int ho = a.hi - 2, wo = a.wi - 2;
        int tile = 0;
        TcUnit t;
        for (int64_t u = blockIdx.x; rn_tc_unit(a, u, t); u += gridDim.x) {
            for (int j = 0; j < t.rows; j++, tile++) {
                const int as = tile & 1;
                rn_tc_mbar_wait(&acc_full[as], (tile >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t v0[32], v1[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 64);
                rn_tc_ld32(taddr, v0);
                rn_tc_ld32(taddr + 32, v1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                rn_tc_mbar_arrive(&acc_empty[as]);                       // the MMA warp may overwrite this stage
                const int ox = t.x0 + px, oy = t.oy0 + j;
                if (ox < wo && oy < ho) {
                    const int64_t o = (((int64_t)t.img * ho + oy) * wo + ox) * 32;
                    float4 *dh = reinterpret_cast<float4 *>(a.out_hi + o);
                    float4 *dl = a.out_lo ? reinterpret_cast<float4 *>(a.out_lo + o) : nullptr;
#pragma unroll
                    for (int c4 = 0; c4 < 8; c4++) {
                        float r[4], l[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int c = c4 * 4 + e;
                            float y = fmaf(__uint_as_float(v0[c]) + __uint_as_float(v1[c]), sScale[c], sShift[c]);
                            if (a.relu) y = fmaxf(y, 0.f);
                            if (dl) {
                                const float h = __uint_as_float(__float_as_uint(y) & 0xffffe000u);
                                l[e] = y - h;
                                y = h;
                            }
                            r[e] = y;
                        }
                        dh[c4] = make_float4(r[0], r[1], r[2], r[3]);
                        if (dl) dl[c4] = make_float4(l[0], l[1], l[2], l[3]);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}
