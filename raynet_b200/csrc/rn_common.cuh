// rn_common.cuh -- shared device helpers for the sm_100a RayNet hot path.
//
// Compiled with -fmad=false: every `a*b+c` below is a separate IEEE multiply and add,
// exactly like the CPU oracle (gcc -ffp-contract=off).  That is what makes the integer
// decisions of the path (voxel indices, feature-map pixels, plane brackets) bit-exact.
// Where a fused multiply-add is wanted for speed and the result is tolerance-gated, it
// is requested explicitly with fmaf().
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "raynet_b200.h"

#define RN_FULL_MASK 0xffffffffu
#define RN_VOX_PER_LANE 4                       // consecutive voxels owned by one lane
#define RN_CHUNK (32 * RN_VOX_PER_LANE)         // voxels one warp covers per chunk (128)
#define RN_MAX_NCH 12                           // longest resident ray = 12 * 128 = 1536 voxels (C5: 3 x 512)

// Device-side parameter block derived from RnParams on the host (rn_api.cu: make_dev).
struct RnDev {
    int M, D, V, F, H, W, pad;
    int gx, gy, gz;
    float bbox[6];
    float bin[3];       // (max - min) / grid in f32, ray_tracing.pyx:103-104
    int fh, fw;         // feature-map extents H+p+1, W+p+1 (feature_similarities.cu:74-75)
    int shift;          // padding - (padding-1)/2     (feature_similarities.cu:49-50)
    int npairs;         // (V*(V-1))/2                 (feature_similarities.cu:106)
    int code_stride;    // bytes per ray in the step-code array (32 per 128-voxel chunk)
    int row_stride;     // floats per ray in S / msgs rows (= M for the reference layout)
    // bricked accumulator layout (rn_engine.cuh): 4x4x2-voxel lines of four 2x2x2 sectors
    int bbx, bby, blz;  // bricks along x, bricks along y, lines along z (= ceil(G/4), ceil(G/4), ceil(G/2))
    int bsx, bsy;       // element strides of one brick step along x / y (bby*blz*32, blz*32)
};

__device__ __forceinline__ float rn_clampf(float x, float a, float b) {
    return fminf(fmaxf(x, a), b);   // cuda_implementations/utils.cu:1-3
}

// ---- cache-hinted memory operations ---------------------------------------------------
// Per-ray rows (s_hat, messages, step codes) are streamed once per sweep: keep them out
// of L1 and mark them evict-first in L2 so the two accumulator grids (gathered and
// scatter-added by every ray) stay L2-resident.
// On sm_100 the .L2::evict_* priority qualifiers are only accepted on 256-bit accesses, so
// the hints travel as createpolicy descriptors (.L2::cache_hint) instead.
__device__ __forceinline__ uint64_t rn_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t rn_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 rn_ld_stream4(const float *p) {
    float4 v;
    uint64_t pol = rn_policy_evict_first();
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void rn_st_stream4(float *p, float4 v) {
    uint64_t pol = rn_policy_evict_first();
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ uint32_t rn_ld_stream_u8(const uint8_t *p) {
    uint32_t v;
    uint64_t pol = rn_policy_evict_first();
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
// Accumulator gather: read-only for the whole kernel, keep in L1 and prefer to keep in L2.
__device__ __forceinline__ float rn_ld_acc(const float *p) {
    float v;
    uint64_t pol = rn_policy_evict_last();
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
// Fire-and-forget scatter-add (RED, no return value), resolved in L2.
__device__ __forceinline__ void rn_red_add(float *p, float v) {
    uint64_t pol = rn_policy_evict_last();
    asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(p), "f"(v), "l"(pol) : "memory");
}
// The same two with a caller-held policy descriptor (created once per kernel).
#ifndef RN_POL_ACC
#define RN_POL_ACC 3      // bit 0: evict-last hint on the accumulator gathers, bit 1: on the scatter-adds
#endif
__device__ __forceinline__ float rn_ld_acc_pol(const float *p, uint64_t pol) {
    float v;
    if (RN_POL_ACC & 1) asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    else asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void rn_red_add_pol(float *p, float v, uint64_t pol) {
    if (RN_POL_ACC & 2) asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(p), "f"(v), "l"(pol) : "memory");
    else asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

// ---- warp scans ------------------------------------------------------------------------
__device__ __forceinline__ float rn_warp_incl_scan_add(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float t = __shfl_up_sync(RN_FULL_MASK, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ float rn_warp_incl_scan_mul(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float t = __shfl_up_sync(RN_FULL_MASK, v, d);
        if (lane >= d) v *= t;
    }
    return v;
}
// inclusive suffix sum: lane l gets sum over lanes >= l
__device__ __forceinline__ float rn_warp_incl_rscan_add(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float t = __shfl_down_sync(RN_FULL_MASK, v, d);
        if (lane + d < 32) v += t;
    }
    return v;
}
__device__ __forceinline__ uint32_t rn_warp_incl_scan_u32(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(RN_FULL_MASK, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ float rn_warp_sum(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(RN_FULL_MASK, v, d);
    return v;
}
__device__ __forceinline__ float rn_warp_max(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v = fmaxf(v, __shfl_xor_sync(RN_FULL_MASK, v, d));
    return v;
}

// ---- a1: sample_in_bbox (sampling_schemes.cu:5-90), operation for operation ------------
__device__ __forceinline__ void rn_sample_in_bbox(int ray_idx, const RnDev &p, const float *Pinv,
                                                  const float *C, float *rs, float *re) {
    float px = (float)(ray_idx / p.H);
    float py = (float)(ray_idx % p.H);
    double out[3], nrm;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        double acc = 0.0;
        acc += (double)(Pinv[r * 3 + 0] * px);
        acc += (double)(Pinv[r * 3 + 1] * py);
        acc += (double)Pinv[r * 3 + 2] * 1.0;
        out[r] = acc;
    }
    nrm = 0.0;
    nrm += (double)(Pinv[9] * px);
    nrm += (double)(Pinv[10] * py);
    nrm += (double)Pinv[11] * 1.0;
    float dir[3];
#pragma unroll
    for (int i = 0; i < 3; i++) dir[i] = (float)(out[i] / nrm - (double)C[i]);
    float t_near = -INFINITY, t_far = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float t1 = (float)(((double)p.bbox[a] - (double)C[a]) / (double)dir[a]);
        float t2 = (float)(((double)p.bbox[3 + a] - (double)C[a]) / (double)dir[a]);
        t_near = fmaxf(fminf(t1, t2), t_near);
        t_far = fminf(fmaxf(t1, t2), t_far);
    }
    float near_mask = (fabsf(t_near) < fabsf(t_far)) ? 1.0f : 0.0f;
    float tn = t_near * near_mask + t_far * (1 - near_mask);
    float tf = (1 - near_mask) * t_near + near_mask * t_far;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        rs[i] = C[i] + tn * dir[i];
        re[i] = C[i] + tf * dir[i];
    }
}

// ---- a2 helpers: projection + pixel -> feature-map element offset ----------------------
// feature_similarities.cu:10-32 (dot_m34v3) and :42-61 (pixel_to_features).
__device__ __forceinline__ int rn_project_offset(const RnDev &p, const float *Pm /*12*/, int view,
                                                 const float *pt) {
    float o0 = 0.f, o1 = 0.f, nz = 0.f;
    o0 += Pm[0] * pt[0]; o0 += Pm[1] * pt[1]; o0 += Pm[2] * pt[2]; o0 += Pm[3] * 1;
    o1 += Pm[4] * pt[0]; o1 += Pm[5] * pt[1]; o1 += Pm[6] * pt[2]; o1 += Pm[7] * 1;
    nz += Pm[8] * pt[0]; nz += Pm[9] * pt[1]; nz += Pm[10] * pt[2]; nz += Pm[11] * 1;
    o0 /= nz;
    o1 /= nz;
    int fx = (int)(roundf(o0) + (float)p.shift);
    int fy = (int)(roundf(o1) + (float)p.shift);
    fx = max(fx, 0); fx = min(fx, p.W);
    fy = max(fy, 0); fy = min(fy, p.H);
    if (fx == 0 || fy == 0) fx = fy = 0;
    return ((view * p.fh + fy) * p.fw + fx) * p.F;
}

// ---- linear voxel index: Gy*Gz*x + Gz*y + z  (mrf_bp.cu:3-10) --------------------------
__device__ __forceinline__ int rn_lin(const RnDev &p, int x, int y, int z) {
    return (x * p.gy + y) * p.gz + z;
}

// ---- occupancy-to-ray message (mrf_bp.cu:12-35 == mrf_np.py:52-71) ---------------------
// o = clamp(sigmoid(acc - msg), 1e-4, 1-1e-4) with the max-shifted exponentials:
//   x >= 0: t2 = 1, t1 = e^-x ;  x < 0: t1 = 1, t2 = e^x  ->  one exp of -|x|.
__device__ __forceinline__ float rn_occ_to_ray(float acc, float msg) {
    float x = acc - msg;
    float e = expf(-fabsf(x));
    float num = (x >= 0.f) ? 1.0f : e;
    float o = num / (1.0f + e);
    return rn_clampf(o, 1e-4f, 0.9999f);   // (float)1e-4, (float)(1-1e-4)
}
