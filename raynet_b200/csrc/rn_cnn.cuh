// rn_cnn.cuh -- SURVEY.md 8(f) row 1: the MV-CNN feature extractor in front of the hot path.
//
// Reference: models.py:90-111 (create_simple_cnn): 5 x [Conv2D(32, 3x3, 'valid') + BatchNormalization]
// with ReLU after the first four, applied to views zero-padded by `padding` = 11 pixels
// (forward_pass.py:181-198), so that an (H, W) image yields the (H + 12, W + 12, 32) channels-last
// feature map the similarity kernel gathers from.  Inference-mode batch normalisation is an affine
// map per channel; it is folded together with the convolution bias into (scale, shift):
//     y = scale[c] * conv(x)[c] + shift[c],   scale = gamma / sqrt(var + eps),
//     shift = beta + scale * (bias - mean).
//
// One layer = one launch of conv3x3_kernel<CIN>: direct convolution in fp32 on the CUDA cores
// (fp32 accumulation is what keeps the features inside the 1e-5 parity budget of the BP marginals;
// single-pass TF32 / BF16 tensor-core products do not).  A CTA of 128 threads owns an 8 x 32 pixel
// output tile for all 32 output channels: the (10 x 34 x CIN) input tile and the 3x3xCINx32 weights
// sit in shared memory, a thread accumulates 8 pixels x 8 output channels in registers (64 FMAs for
// every 10 + 6 shared-memory loads of a (ky, cin) step, so the FMA pipe, not the LSU, is the limiter),
// the shared-memory layout (channel swizzle for CIN = 32, odd pixel stride otherwise) sends the 8 pixel
// addresses of a warp to 8 banks; the input tile of the 32-channel layers arrives by 16-byte cp.async.
// CTAs are persistent (grid = 2 per SM) and keep the weights resident while they walk the tiles.
#pragma once

#include "rn_common.cuh"

#define RN_CNN_COUT 32
#define RN_CNN_TH 8
#define RN_CNN_TW 32

__device__ __forceinline__ uint64_t rn_cnn_pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void rn_cnn_unpack2(uint64_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t rn_cnn_fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// channel swizzle of pixel (py, px) of the shared input tile: a multiple of 4 (16-byte groups intact),
// different for the 8 pixels (4 column blocks x 2 rows) one warp reads in one instruction
__device__ __forceinline__ int rn_cnn_swz(int py, int px) { return (((px >> 3) & 3) << 3) | ((py & 1) << 2); }
__device__ __forceinline__ void rn_cnn_cp_async16(float *dst_smem, const float *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

struct ConvArgs {
    const float *in;       // [N][Hi][Wi][CIN]
    const float *w;        // [3][3][CIN][32]   (Keras kernel layout)
    const float *scale;    // [32]
    const float *shift;    // [32]
    float *out;            // [N][Hi-2][Wi-2][32]
    int n, hi, wi, relu;
    float *out_lo;         // optional: out receives the TF32-exact high part of the result, out_lo the rest (rn_cnn_tc.cuh)
};

template <int CIN>
__host__ __device__ constexpr int rn_cnn_smem_words() {
    return 9 * CIN * RN_CNN_COUT + (RN_CNN_TH + 2) * (RN_CNN_TW + 2) * ((CIN == 32) ? 32 : ((CIN % 2 == 0) ? CIN + 1 : CIN)) + 2 * RN_CNN_COUT + 4;
}

template <int CIN>
__global__ void __launch_bounds__(128) conv3x3_kernel(ConvArgs a) {
    extern __shared__ __align__(16) float cnn_smem[];
    // CIN = 32: pixels are 32 words apart and the channel index is XOR-swizzled with the pixel position
    // (rn_cnn_swz) -- 16-byte groups stay contiguous for cp.async and the 8 pixel addresses of a warp
    // still hit 8 banks; other CIN: an odd pixel stride does the same job
    constexpr bool kSwz = (CIN == 32);
    constexpr int PS = kSwz ? 32 : ((CIN % 2 == 0) ? CIN + 1 : CIN);
    constexpr int TWI = RN_CNN_TW + 2, THI = RN_CNN_TH + 2;
    float *sW = cnn_smem;                             // [9][CIN][32]
    float *sIn = sW + 9 * CIN * RN_CNN_COUT;          // [THI][TWI][PS]
    float *sScale = sIn + ((THI * TWI * PS + 3) & ~3);
    float *sShift = sScale + RN_CNN_COUT;
    const int tid = threadIdx.x;
    for (int i = tid; i < 9 * CIN * RN_CNN_COUT; i += 128) sW[i] = __ldg(a.w + i);
    if (tid < RN_CNN_COUT) { sScale[tid] = __ldg(a.scale + tid); sShift[tid] = __ldg(a.shift + tid); }

    const int ho = a.hi - 2, wo = a.wi - 2;
    const int tiles_x = (wo + RN_CNN_TW - 1) / RN_CNN_TW, tiles_y = (ho + RN_CNN_TH - 1) / RN_CNN_TH;
    const int64_t n_tiles = (int64_t)a.n * tiles_y * tiles_x;
    const int qg = tid & 3;            // output channels 8 qg .. 8 qg + 7
    const int pg = tid >> 2;           // pixel group: row pr, columns 8 pc .. 8 pc + 7 of the tile
    const int pr = pg >> 2, pc = pg & 3;

    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int img = (int)(t / (tiles_y * tiles_x));
        const int rem = (int)(t - (int64_t)img * tiles_y * tiles_x);
        const int y0 = (rem / tiles_x) * RN_CNN_TH, x0 = (rem % tiles_x) * RN_CNN_TW;
        __syncthreads();   // the previous tile's reads of sIn are done (and the weights are in place)
        // ---- input tile: rows y0 .. y0+9, columns x0 .. x0+33, zero outside the image ---------------
        const float *src = a.in + (int64_t)img * a.hi * a.wi * CIN;
        if constexpr (kSwz) {
            for (int i = tid; i < THI * TWI * (CIN / 4); i += 128) {      // one 16-byte group of channels per copy
                const int g4 = i % (CIN / 4), px = (i / (CIN / 4)) % TWI, py = i / ((CIN / 4) * TWI);
                const int gy = y0 + py, gx = x0 + px;
                float *dst = sIn + (py * TWI + px) * PS + ((4 * g4) ^ rn_cnn_swz(py, px));
                if (gy < a.hi && gx < a.wi) rn_cnn_cp_async16(dst, src + ((int64_t)gy * a.wi + gx) * CIN + 4 * g4);
                else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
        } else {
            for (int i = tid; i < THI * TWI * CIN; i += 128) {
                const int c = i % CIN, px = (i / CIN) % TWI, py = i / (CIN * TWI);
                const int gy = y0 + py, gx = x0 + px;
                float v = 0.f;
                if (gy < a.hi && gx < a.wi) v = __ldg(src + ((int64_t)gy * a.wi + gx) * CIN + c);
                sIn[(py * TWI + px) * PS + c] = v;
            }
        }
        __syncthreads();
        // ---- 8 pixels x 8 channels per thread ----------------------------------------------------------
        // accumulators as packed pairs of output channels: one FFMA2 (fma.rn.f32x2) performs the two
        // FMAs of (pixel i, channels 2 qq and 2 qq + 1) -- half the issue slots of scalar FFMAs
        uint64_t acc2[8][4];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc2[i][q] = 0ull;
#pragma unroll 1
        for (int ky = 0; ky < 3; ky++) {
            const float *row = sIn + ((pr + ky) * TWI + pc * 8) * PS;
            const int sw_lo = kSwz ? rn_cnn_swz(pr + ky, pc * 8) : 0, sw_hi = kSwz ? rn_cnn_swz(pr + ky, pc * 8 + 8) : 0;
#pragma unroll 2
            for (int c = 0; c < CIN; c++) {
                uint64_t v2[10];
#pragma unroll
                for (int i = 0; i < 10; i++) {
                    const float v = row[i * PS + (c ^ (i < 8 ? sw_lo : sw_hi))];
                    v2[i] = rn_cnn_pack2(v, v);
                }
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    const ulonglong2 wa = *reinterpret_cast<const ulonglong2 *>(sW + ((ky * 3 + kx) * CIN + c) * RN_CNN_COUT + qg * 8);
                    const ulonglong2 wb = *reinterpret_cast<const ulonglong2 *>(sW + ((ky * 3 + kx) * CIN + c) * RN_CNN_COUT + qg * 8 + 4);
                    const uint64_t w2[4] = {wa.x, wa.y, wb.x, wb.y};
#pragma unroll
                    for (int i = 0; i < 8; i++)
#pragma unroll
                        for (int q = 0; q < 4; q++) acc2[i][q] = rn_cnn_fma2(v2[i + kx], w2[q], acc2[i][q]);
                }
            }
        }
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int q = 0; q < 4; q++) rn_cnn_unpack2(acc2[i][q], acc[i][2 * q], acc[i][2 * q + 1]);
        // ---- epilogue: folded batch norm (+ ReLU), channels-last store --------------------------------
        const int oy = y0 + pr;
        if (oy < ho) {
            float sc[8], sh[8];
#pragma unroll
            for (int q = 0; q < 8; q++) { sc[q] = sScale[qg * 8 + q]; sh[q] = sShift[qg * 8 + q]; }
            float *dst = a.out + (((int64_t)img * ho + oy) * wo) * RN_CNN_COUT + qg * 8;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int ox = x0 + pc * 8 + i;
                if (ox < wo) {
                    float r[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        r[q] = fmaf(acc[i][q], sc[q], sh[q]);
                        if (a.relu) r[q] = fmaxf(r[q], 0.f);
                    }
                    if (a.out_lo) {   // hi / lo pair for the tensor-core layers: hi exactly representable in TF32
                        float l[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            const float h = __uint_as_float(__float_as_uint(r[q]) & 0xffffe000u);
                            l[q] = r[q] - h;
                            r[q] = h;
                        }
                        float4 *l4 = reinterpret_cast<float4 *>(a.out_lo + (dst - a.out) + (int64_t)ox * RN_CNN_COUT);
                        l4[0] = make_float4(l[0], l[1], l[2], l[3]);
                        l4[1] = make_float4(l[4], l[5], l[6], l[7]);
                    }
                    float4 *d4 = reinterpret_cast<float4 *>(dst + (int64_t)ox * RN_CNN_COUT);
                    d4[0] = make_float4(r[0], r[1], r[2], r[3]);
                    d4[1] = make_float4(r[4], r[5], r[6], r[7]);
                }
            }
        }
    }
}
