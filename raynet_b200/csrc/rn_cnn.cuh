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
// pixels are padded to CIN + 1 words so that the 8 pixel addresses of a warp fall into 8 banks.
// CTAs are persistent (grid = 2 per SM) and keep the weights resident while they walk the tiles.
#pragma once

#include "rn_common.cuh"

#define RN_CNN_COUT 32
#define RN_CNN_TH 8
#define RN_CNN_TW 32

struct ConvArgs {
    const float *in;       // [N][Hi][Wi][CIN]
    const float *w;        // [3][3][CIN][32]   (Keras kernel layout)
    const float *scale;    // [32]
    const float *shift;    // [32]
    float *out;            // [N][Hi-2][Wi-2][32]
    int n, hi, wi, relu;
};

template <int CIN>
__host__ __device__ constexpr int rn_cnn_smem_words() {
    return 9 * CIN * RN_CNN_COUT + (RN_CNN_TH + 2) * (RN_CNN_TW + 2) * ((CIN % 2 == 0) ? CIN + 1 : CIN) + 2 * RN_CNN_COUT + 4;
}

template <int CIN>
__global__ void __launch_bounds__(128) conv3x3_kernel(ConvArgs a) {
    extern __shared__ __align__(16) float cnn_smem[];
    constexpr int PS = (CIN % 2 == 0) ? CIN + 1 : CIN;   // odd pixel stride: the 8 pixel addresses of a warp hit 8 banks
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
        for (int i = tid; i < THI * TWI * CIN; i += 128) {
            const int c = i % CIN, px = (i / CIN) % TWI, py = i / (CIN * TWI);
            const int gy = y0 + py, gx = x0 + px;
            float v = 0.f;
            if (gy < a.hi && gx < a.wi) v = __ldg(src + ((int64_t)gy * a.wi + gx) * CIN + c);
            sIn[(py * TWI + px) * PS + c] = v;
        }
        __syncthreads();
        // ---- 8 pixels x 8 channels per thread ----------------------------------------------------------
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int q = 0; q < 8; q++) acc[i][q] = 0.f;
#pragma unroll 1
        for (int ky = 0; ky < 3; ky++) {
            const float *row = sIn + ((pr + ky) * TWI + pc * 8) * PS;
#pragma unroll 2
            for (int c = 0; c < CIN; c++) {
                float v[10];
#pragma unroll
                for (int i = 0; i < 10; i++) v[i] = row[i * PS + c];
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    const float4 w0 = *reinterpret_cast<const float4 *>(sW + ((ky * 3 + kx) * CIN + c) * RN_CNN_COUT + qg * 8);
                    const float4 w1 = *reinterpret_cast<const float4 *>(sW + ((ky * 3 + kx) * CIN + c) * RN_CNN_COUT + qg * 8 + 4);
                    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int i = 0; i < 8; i++)
#pragma unroll
                        for (int q = 0; q < 8; q++) acc[i][q] = fmaf(v[i + kx], w[q], acc[i][q]);
                }
            }
        }
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
                    float4 *d4 = reinterpret_cast<float4 *>(dst + (int64_t)ox * RN_CNN_COUT);
                    d4[0] = make_float4(r[0], r[1], r[2], r[3]);
                    d4[1] = make_float4(r[4], r[5], r[6], r[7]);
                }
            }
        }
    }
}
