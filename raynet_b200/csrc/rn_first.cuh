// rn_first.cuh -- the plane->voxel mapping (a4) fused into the FIRST BP sweep (a5 + a6).
//
// Round 1 ran planemap3_kernel (issue-bound: 73 % of the issue slots, 3.7 ms per 9 images on C3) to write the
// `lin` / `s_hat` rows and then bp4_kernel<NCH, true> (memory-bound, 58 % of the issue slots idle, 3.0 ms) to read them
// straight back.  Here the warp that sweeps a ray for the first time builds the ray's rows itself: voxel coordinates
// from the 2-bit step codes (popc), bricked offsets and centres from per-axis tables in shared memory, the plane
// distribution S_planes (256 B per ray, written by simscore3_kernel) interpolated at the voxel centres
// (planes_voxels_mapping.cu:6-92), normalised, clipped and renormalised (mrf_np.py:4-8) -- identical arithmetic to
// planemap3_kernel -- leaves them in shared memory for rn_bp4_ray<NCH, true> and writes them to HBM once for the
// later sweeps.  The first sweep straight after a reset reads a uniform accumulator and no messages, so the fused
// kernel's only inputs are 0.25 B/voxel of codes and 280 B per ray.
#pragma once

#include "rn_bp4.cuh"
#include "rn_simmap3.cuh"

struct FirstArgs {
    const float *axes;            // [Gx+Gy+Gz] voxel-centre coordinates per axis
    const float *starts, *ends;   // [n][3]
    const uint32_t *hdr;          // [n][2]
    const uint8_t *codes;         // [n][code_stride]
    const int32_t *count;         // [n]
    const float *S_planes;        // [n][D]
    int32_t *lin;                 // [n][row_stride] out
    float *s_hat;                 // [n][row_stride] out
    float *msgs;                  // [n][row_stride] out
    const float *acc_in;          // uniform: acc_in[0] is the value of every voxel
    float *acc_out;
    const int32_t *order;
    int64_t first, n;
    int rays_per_warp;
    int map_only;                 // rays BP skips (count <= 1, mrf_np.py:299-301): build their rows, no sweep
};

// dynamic shared memory (words): per CTA the per-axis table, per warp (S_k, S_k+1 - S_k) pairs + sLin[ROW] + sS[ROW]
__host__ __device__ inline size_t rn_first_cta_words(int gsum) { return 2 * (size_t)((gsum + 1) & ~1); }
__host__ __device__ inline size_t rn_first_warp_words(int D, int nch) { return 2 * (size_t)((D + 3) & ~1) + 2 * (size_t)nch * RN_CHUNK; }

template <int NCH>
__global__ void __launch_bounds__(128) bp4_first_mapped_kernel(RnDev p, FirstArgs a) {
    extern __shared__ __align__(16) unsigned char rn_first_smem[];
    constexpr int ROW = NCH * RN_CHUNK;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int D = p.D;
    const int gsum = p.gx + p.gy + p.gz;
    int2 *sTab = reinterpret_cast<int2 *>(rn_first_smem);   // per axis entry: (bricked offset contribution, centre coordinate)
    float *warp0 = reinterpret_cast<float *>(sTab + ((gsum + 1) & ~1));
    float *wbase = warp0 + (size_t)wid * rn_first_warp_words(D, NCH);
    float2 *sS2 = reinterpret_cast<float2 *>(wbase);
    int *sLin = reinterpret_cast<int *>(wbase + 2 * ((D + 3) & ~1));
    float *sS = reinterpret_cast<float *>(sLin + ROW);
    for (int i = threadIdx.x; i < gsum; i += blockDim.x) {
        const int b = (i < p.gx) ? rn_brick_fx(p, i) : (i < p.gx + p.gy) ? rn_brick_fy(p, i - p.gx) : rn_brick_fz(i - p.gx - p.gy);
        sTab[i] = make_int2(b, __float_as_int(__ldg(a.axes + i)));
    }
    __syncthreads();
    const uint64_t pol_stream = rn_policy_evict_first();
    const uint64_t pol_keep = rn_policy_evict_last();
    const float fDm1 = (float)(D - 1);
    const float pstep = (1.0f - 0.0f) / fDm1;

    const int rpw = a.rays_per_warp;
    const int64_t k0 = (int64_t)blockIdx.x * (4 * rpw) + wid;
    for (int t = 0; t < rpw; t++) {
        const int64_t k = k0 + 4 * (int64_t)t;
        if (k >= a.n) break;
        const int64_t r = a.order ? (int64_t)__ldg(a.order + a.first + k) : a.first + k;
        const int L = __ldg(a.count + r);
        float rs[3], re[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            rs[i] = __ldg(a.starts + 3 * r + i);
            re[i] = __ldg(a.ends + 3 * r + i);
        }
        __syncwarp();   // the previous ray's reads of this warp's buffers are done
        for (int kk = lane; kk < D; kk += 32) {
            const float s0 = __ldg(a.S_planes + r * (int64_t)D + kk);
            const float s1 = (kk + 1 < D) ? __ldg(a.S_planes + r * (int64_t)D + kk + 1) : s0;
            sS2[kk] = make_float2(s0, s1 - s0);
        }
        __syncwarp();
        // ---- plane -> voxel mapping: exactly planemap3_kernel's arithmetic -------------------------------------
        float ray[3];
#pragma unroll
        for (int i = 0; i < 3; i++) ray[i] = re[i] - rs[i];
        float ray_norm = 0.f;
#pragma unroll
        for (int i = 0; i < 3; i++) ray_norm += ray[i] * ray[i];
        const float rayn[3] = {ray[0] / ray_norm, ray[1] / ray_norm, ray[2] / ray_norm};
        const RayHead head = rn_ray_head(a.hdr + 2 * r);
        const uint2 *words = reinterpret_cast<const uint2 *>(a.codes + r * (int64_t)p.code_stride);
        int32_t *lin_row = a.lin + r * (int64_t)p.row_stride;
        StepCount before = {0, 0, 0};
        float lsum = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const uint4 cwA = __ldg(reinterpret_cast<const uint4 *>(words + c * 4));
            const uint4 cwB = __ldg(reinterpret_cast<const uint4 *>(words + c * 4 + 2));
            const uint32_t lo[4] = {cwA.x, cwA.z, cwB.x, cwB.z}, hi[4] = {cwA.y, cwA.w, cwB.y, cwB.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = c * RN_CHUNK + 32 * j + lane;
                int vx, vy, vz;
                rn_decode_pair(head, lo[j], hi[j], lane, before, vx, vy, vz);
                float out = 0.f;
                if (c < NCH - 1 || i < L) {
                    const int2 ex = sTab[vx], ey = sTab[p.gx + vy], ez = sTab[p.gx + p.gy + vz];
                    const int lin = ex.x + ey.x + ez.x;
                    sLin[i] = lin;
                    lin_row[i] = lin;
                    const float sum = (__int_as_float(ex.y) - rs[0]) * rayn[0] + (__int_as_float(ey.y) - rs[1]) * rayn[1] +
                                      (__int_as_float(ez.y) - rs[2]) * rayn[2];
                    const float tt = rn_clampf(sum, 1e-4f, 1 - 1e-4f);
                    const int left = min((int)(tt * fDm1), D - 2);
                    const float2 sd = sS2[left];
                    out = fmaf((tt - (float)left * pstep) * fDm1, sd.y, sd.x);
                    lsum += out;
                }
                sS[i] = out;
            }
        }
        const float inv_sr = 1.0f / rn_warp_sum(lsum);
        __syncwarp();
        // normalise; clip + renormalise (mrf_np.py:4-8); each lane owns the quads the scans will read
        float csum = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const int i = c * RN_CHUNK + 4 * lane;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < NCH - 1 || i < L) {
                v = *reinterpret_cast<const float4 *>(sS + i);
                v.x = rn_clampf(v.x * inv_sr, 1e-5f, 0.99999f);
                v.y = (c < NCH - 1 || i + 1 < L) ? rn_clampf(v.y * inv_sr, 1e-5f, 0.99999f) : 0.f;
                v.z = (c < NCH - 1 || i + 2 < L) ? rn_clampf(v.z * inv_sr, 1e-5f, 0.99999f) : 0.f;
                v.w = (c < NCH - 1 || i + 3 < L) ? rn_clampf(v.w * inv_sr, 1e-5f, 0.99999f) : 0.f;
                csum += (v.x + v.y) + (v.z + v.w);
            }
            *reinterpret_cast<float4 *>(sS + i) = v;
        }
        const float inv_c = 1.0f / rn_warp_sum(csum);
        float *out_row = a.s_hat + r * (int64_t)p.row_stride;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const int i = c * RN_CHUNK + 4 * lane;
            float4 v = *reinterpret_cast<const float4 *>(sS + i);
            v.x *= inv_c; v.y *= inv_c; v.z *= inv_c; v.w *= inv_c;
            *reinterpret_cast<float4 *>(sS + i) = v;
            if (c < NCH - 1 || i < L) rn_st_stream4(out_row + i, v);
        }
        __syncwarp();
        // ---- the first sweep on the rows just built ----------------------------------------------------------------
        if (!a.map_only)
            rn_bp4_ray<NCH, true>(a.acc_in, a.acc_out, true, sLin, sS, nullptr, a.msgs + r * (int64_t)p.row_stride, L, lane,
                                  pol_stream, pol_keep);
    }
}
