// rn_kernels.cuh -- front-end and reference-layout kernels of the RayNet hot path (sm_100a).
//
//   dda_kernel      thread-per-ray Amanatides-Woo traversal (a3), optionally fused with
//                   sample_in_bbox (a1), emitting the reference's int32 [M][3] lists (the
//                   resident step-code flavour lives in rn_engine.cuh).
//   simmap_kernel   warp-per-ray: sample_in_bbox (a1) -> plane-sweep similarity + softmax
//                   (a2) -> plane->voxel interpolation (a4) -> clip_and_renorm; serves both
//                   the reference layout (kAos) and the resident pipeline.
//   small stand-alone kernels (sample points, planes->voxels on given S, fills, ...).
// The BP sweep and the depth pass (both layouts) are bp2_kernel / depth2_kernel in
// rn_engine.cuh.
#pragma once

#include "rn_common.cuh"
#include "rn_engine.cuh"

// =======================================================================================
// a3. DDA  (ray_tracing.pyx:99-199 == ray_tracing.cu:15-142), thread per ray
// =======================================================================================
__device__ __forceinline__ int rn_dda(const RnDev &p, const float *rs_in, const float *re_in, int32_t *idx_row) {
    const float EPS = 1e-2f;
    float s[3], e[3], ray[3], tMax[3], tDelta[3];
    int step[3], cur[3], last[3];
    const int g[3] = {p.gx, p.gy, p.gz};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        s[a] = rs_in[a] - p.bbox[a];
        e[a] = re_in[a] - p.bbox[a];
        ray[a] = e[a] - s[a];
        step[a] = ray[a] >= 0 ? 1 : -1;
        float nudge = ((float)step[a] * p.bin[a]) * EPS;
        s[a] = s[a] + nudge;
        e[a] = e[a] - nudge;
        cur[a] = (int)floorf(s[a] / p.bin[a]);
        last[a] = (int)floorf(e[a] / p.bin[a]);
    }
    bool inside = cur[0] >= 0 && cur[0] < g[0] && cur[1] >= 0 && cur[1] < g[1] && cur[2] >= 0 && cur[2] < g[2];
    if (!inside) return 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        tMax[a] = FLT_MAX;
        tDelta[a] = FLT_MAX;
        if (ray[a] != 0) {
            float cc = (float)cur[a] * p.bin[a];
            float b = (step[a] < 0 && cc < s[a]) ? cc : cc + (float)step[a] * p.bin[a];
            tMax[a] = (b - s[a]) / ray[a];
            tDelta[a] = ((float)step[a] * p.bin[a]) / ray[a];
        }
    }
    idx_row[0] = cur[0];
    idx_row[1] = cur[1];
    idx_row[2] = cur[2];
    int ii = 1;
    const int M = p.M;
    while (!(cur[0] == last[0] && cur[1] == last[1] && cur[2] == last[2]) && ii < M) {
        // strict '<', ties X=Y go to the Y/Z branch, any tie with Z goes to Z
        bool xy = tMax[0] < tMax[1];
        float tm = xy ? tMax[0] : tMax[1];
        int a = xy ? 0 : 1;
        a = (tm < tMax[2]) ? a : 2;
        bool ax = (a == 0), ay = (a == 1), az = (a == 2);
        cur[0] += ax ? step[0] : 0;
        cur[1] += ay ? step[1] : 0;
        cur[2] += az ? step[2] : 0;
        int ca = ax ? cur[0] : (ay ? cur[1] : cur[2]);
        int ga = ax ? g[0] : (ay ? g[1] : g[2]);
        if (ca < 0 || ca >= ga) break;
        tMax[0] = ax ? tMax[0] + tDelta[0] : tMax[0];
        tMax[1] = ay ? tMax[1] + tDelta[1] : tMax[1];
        tMax[2] = az ? tMax[2] + tDelta[2] : tMax[2];
        idx_row[3 * ii + 0] = cur[0];
        idx_row[3 * ii + 1] = cur[1];
        idx_row[3 * ii + 2] = cur[2];
        ii++;
    }
    return ii;
}

struct DdaArgs {
    const int32_t *ray_idxs;   // if non-null: start/end come from sample_in_bbox (and are written out)
    const float *P_inv, *centre;
    float *starts, *ends;      // inputs when ray_idxs == null; optional outputs otherwise
    int32_t *idx;              // [n][M][3]
    int32_t *count;            // [n]
    int64_t n_rays;
};

__global__ void __launch_bounds__(128) dda_kernel(RnDev p, DdaArgs a) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_rays) return;
    float rs[3], re[3];
    if (a.ray_idxs) {
        float Pinv[12], C[3];
#pragma unroll
        for (int i = 0; i < 12; i++) Pinv[i] = __ldg(a.P_inv + i);
#pragma unroll
        for (int i = 0; i < 3; i++) C[i] = __ldg(a.centre + i);
        rn_sample_in_bbox(__ldg(a.ray_idxs + r), p, Pinv, C, rs, re);
        if (a.starts) {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                a.starts[3 * r + i] = rs[i];
                a.ends[3 * r + i] = re[i];
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            rs[i] = a.starts[3 * r + i];
            re[i] = a.ends[3 * r + i];
        }
    }
    a.count[r] = rn_dda(p, rs, re, a.idx + r * (int64_t)p.M * 3);
}

// =======================================================================================
// a2 + a4. similarity + plane->voxel mapping, warp per ray
// =======================================================================================
struct SimMapArgs {
    const int32_t *ray_idxs;   // non-null: rays start from sample_in_bbox; null: starts/ends are inputs
    const float *features, *P, *P_inv, *centre;
    const int32_t *view_ids;   // optional [V]: slot of each view inside `features` (null = 0..V-1)
    const float *starts_in, *ends_in;   // optional [n][3]: precomputed sample_in_bbox results (same arithmetic)
    float *S_planes;           // [n][D]     optional out
    float *points;             // [n][D][4]  optional out
    float *depth_planes;       // [n]        optional out: |point[argmax_k S] - C|
    // mapping stage (count != null)
    const float *axes;         // [Gx+Gy+Gz] voxel-centre coordinates per axis
    const int32_t *idx;        // kAos
    const uint32_t *hdr;       // !kAos
    const uint8_t *codes;      // !kAos
    const int32_t *count;
    float *S_vox;              // [n][row_stride] optional out: normalised S_voxel_space
    float *s_hat;              // [n][row_stride] optional out: clip_and_renorm(S_voxel_space)
    int32_t *lin;              // [n][row_stride] optional out (!kAos): bricked accumulator offset of every voxel
    float *depth_vox;          // [n] optional out: |centre(argmax voxel of S_vox) - C|
    int64_t n_rays;
    int64_t tile_len;          // > 0: warps walk the rays in 8x8-pixel tiles (rn_tiled_position)
    int val_stride;            // floats of the per-warp voxel buffer (0 without mapping stage)
    int tile_mode;
    int k_lo, k_hi;            // simscore3_kernel: planes [k_lo, k_hi) of this pass (k_hi == 0: all D planes)
    int raw_scores;            // simscore3_kernel: 1 = store the plane scores of the pass, softmax by softmax_planes_kernel
};

// dynamic shared memory: per CTA [V*12 P][12 P_inv][4 C][V view slots], per warp
// [D*V feature offsets][D plane scores][val_stride voxel values]
__host__ __device__ inline size_t rn_simmap_smem_bytes(int D, int V, int val_stride, int warps) {
    return sizeof(float) * (size_t)(V * 12 + 16 + ((V + 3) & ~3)) +
           (size_t)warps * sizeof(float) * ((size_t)((D * V + 3) & ~3) + (size_t)((D + 3) & ~3) + (size_t)val_stride);
}

// One warp per ray.
//  (1) every (plane, view) sample of the ray is projected by one lane (reference arithmetic,
//      IEEE division: the rounded pixel must equal the oracle's) -> feature offsets in smem;
//  (2) plane scores: S_k = sum_{i<j} <f_i, f_j> = 1/2 (|sum_v f_v|^2 - sum_v |f_v|^2).  For F = 32
//      a lane owns 4 channels (one LDG.128) and 8 lanes cover a feature vector, so one warp
//      instruction fetches the vectors of 4 planes; other F fall back to lane = channel;
//  (3) softmax over planes;  (4) plane -> voxel interpolation with the voxel values parked in
//      shared memory between the three passes (sum, clip + renormalise, store).
template <bool kAos>
__global__ void __launch_bounds__(128) simmap_kernel(RnDev p, SimMapArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warps = blockDim.x >> 5;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.D, V = p.V, DV = p.D * p.V;
    float *sP = reinterpret_cast<float *>(smem_raw);
    float *sPinv = sP + V * 12;
    float *sC = sPinv + 12;
    int *sView = reinterpret_cast<int *>(sC + 4);
    float *warp0 = reinterpret_cast<float *>(sView + ((V + 3) & ~3));
    const size_t per_warp = (size_t)((DV + 3) & ~3) + (size_t)((D + 3) & ~3) + (size_t)a.val_stride;
    int *sOff = reinterpret_cast<int *>(warp0 + (size_t)wid * per_warp);
    float *sS = reinterpret_cast<float *>(sOff + ((DV + 3) & ~3));
    float *sVal = sS + ((D + 3) & ~3);

    for (int i = threadIdx.x; i < V * 12; i += blockDim.x) sP[i] = __ldg(a.P + i);
    if (threadIdx.x < 12) sPinv[threadIdx.x] = a.P_inv ? __ldg(a.P_inv + threadIdx.x) : 0.f;
    if (threadIdx.x < 3) sC[threadIdx.x] = a.centre ? __ldg(a.centre + threadIdx.x) : 0.f;
    if (threadIdx.x < V) sView[threadIdx.x] = a.view_ids ? __ldg(a.view_ids + threadIdx.x) : (int)threadIdx.x;
    __syncthreads();

    const int64_t t = (int64_t)blockIdx.x * warps + wid;
    if (t >= a.n_rays) return;
    const int64_t r = rn_tiled_position(t, a.tile_len, p.H, a.tile_mode);

    // ---- a1: ray start / end -------------------------------------------------------
    float rs[3], re[3];
    if (a.starts_in) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            rs[i] = __ldg(a.starts_in + 3 * r + i);
            re[i] = __ldg(a.ends_in + 3 * r + i);
        }
    } else {
        rn_sample_in_bbox(__ldg(a.ray_idxs + r), p, sPinv, sC, rs, re);
    }

    // ---- a2 step 1: project every (plane, view) sample, one sample per lane ----------
    for (int s0 = 0; s0 < DV; s0 += 32) {
        const int sidx = s0 + lane;
        if (sidx < DV) {
            const int k = sidx / V, v = sidx - k * V;
            float pt[3];
#pragma unroll
            for (int i = 0; i < 3; i++) pt[i] = rs[i] + (float)k * (re[i] - rs[i]) / (float)(D - 1);
            sOff[sidx] = rn_project_offset(p, sP + v * 12, sView[v], pt);
        }
    }
    __syncwarp();

    // ---- a2 step 2: plane scores --------------------------------------------------------
    const float inv_pairs = 0.5f / (float)p.npairs;
    if (p.F == 32) {
        const int g = lane >> 3, cl = (lane & 7) * 4;
        for (int k0 = 0; k0 < D; k0 += 4) {
            const int k = min(k0 + g, D - 1);
            const int *offk = sOff + k * V;
            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
            float sq = 0.f;
#pragma unroll 3
            for (int v = 0; v < V; v++) {
                const float4 f = __ldg(reinterpret_cast<const float4 *>(a.features + (int64_t)offk[v] + cl));
                sum.x += f.x; sum.y += f.y; sum.z += f.z; sum.w += f.w;
                sq = fmaf(f.x, f.x, sq); sq = fmaf(f.y, f.y, sq); sq = fmaf(f.z, f.z, sq); sq = fmaf(f.w, f.w, sq);
            }
            float val = fmaf(sum.x, sum.x, fmaf(sum.y, sum.y, fmaf(sum.z, sum.z, fmaf(sum.w, sum.w, -sq))));
            val += __shfl_xor_sync(RN_FULL_MASK, val, 1);
            val += __shfl_xor_sync(RN_FULL_MASK, val, 2);
            val += __shfl_xor_sync(RN_FULL_MASK, val, 4);
            if ((lane & 7) == 0 && k0 + g < D) sS[k0 + g] = val * inv_pairs;
        }
    } else {   // any F: lane = channel (strided), one warp reduction per plane
        for (int k = 0; k < D; k++) {
            const int *offk = sOff + k * V;
            float part = 0.f;
            for (int c0 = 0; c0 < p.F; c0 += 32) {
                const int ch = c0 + lane;
                float sum = 0.f, sq = 0.f;
                if (ch < p.F) {
                    for (int v = 0; v < V; v++) {
                        const float f = __ldg(a.features + (int64_t)offk[v] + ch);
                        sum += f;
                        sq = fmaf(f, f, sq);
                    }
                }
                part += fmaf(sum, sum, -sq);
            }
            part = rn_warp_sum(part);
            if (lane == 0) sS[k] = part * inv_pairs;
        }
    }
    __syncwarp();

    // ---- softmax over the D planes (feature_similarities.cu:109-123) -------------------
    float mx = -INFINITY;
    for (int k = lane; k < D; k += 32) mx = fmaxf(mx, sS[k]);
    mx = rn_warp_max(mx);
    float ssum = 0.f;
    for (int k = lane; k < D; k += 32) {
        const float ev = expf(sS[k] - mx);
        sS[k] = ev;
        ssum += ev;
    }
    ssum = rn_warp_sum(ssum);
    float bv = -INFINITY;   // first arg-max over planes (similarities.py:213-229)
    int bk = 0;
    for (int k = lane; k < D; k += 32) {
        const float v = sS[k] / ssum;
        sS[k] = v;
        if (a.S_planes) a.S_planes[r * (int64_t)D + k] = v;
        if (v > bv) { bv = v; bk = k; }
    }
    __syncwarp();

    if (a.points) {   // similarities.py:206-209 / sampling_schemes.cu:116-121
        for (int k = lane; k < D; k += 32) {
            float4 q;
            q.x = rs[0] + (float)k * (re[0] - rs[0]) / (float)(D - 1);
            q.y = rs[1] + (float)k * (re[1] - rs[1]) / (float)(D - 1);
            q.z = rs[2] + (float)k * (re[2] - rs[2]) / (float)(D - 1);
            q.w = 1.0f;
            reinterpret_cast<float4 *>(a.points)[r * (int64_t)D + k] = q;
        }
    }
    if (a.depth_planes) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            float ov = __shfl_xor_sync(RN_FULL_MASK, bv, d);
            int ok = __shfl_xor_sync(RN_FULL_MASK, bk, d);
            if (ov > bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
        }
        if (lane == 0) {
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                float pt = rs[i] + (float)bk * (re[i] - rs[i]) / (float)(D - 1);
                float dd = pt - sC[i];
                sum += dd * dd;
            }
            a.depth_planes[r] = sqrtf(sum);
        }
    }

    // ---- a4: plane -> voxel mapping (planes_voxels_mapping.cu:6-92) ---------------------
    if (a.count == nullptr) return;
    const int L = __ldg(a.count + r);
    if (L <= 0) {
        if (a.depth_vox && lane == 0) {
            // raynet/mvcnn depth kernels read slot 0 of a zero-filled list: voxel (0,0,0)
            float sum = 0.f;
            float cc[3] = {__ldg(a.axes), __ldg(a.axes + p.gx), __ldg(a.axes + p.gx + p.gy)};
#pragma unroll
            for (int i = 0; i < 3; i++) { float dd = cc[i] - sC[i]; sum += dd * dd; }
            a.depth_vox[r] = sqrtf(sum);
        }
        return;
    }
    const int nch = (L + RN_CHUNK - 1) / RN_CHUNK;
    float ray[3];
#pragma unroll
    for (int i = 0; i < 3; i++) ray[i] = re[i] - rs[i];
    float ray_norm = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) ray_norm += ray[i] * ray[i];
    const float pstep = (1.0f - 0.0f) / (float)(D - 1);
    const float fDm1 = (float)(D - 1);

    // voxel (c, j) of this lane = c*128 + 32*j + lane: lanes walk CONSECUTIVE voxels; the step
    // codes decode with popc (rn_engine.cuh)
    RayHead head = {0, 0, 0, 1, 1, 1};
    const uint2 *words = nullptr;
    const int32_t *idx_row = nullptr;
    if (kAos) {
        idx_row = a.idx + r * (int64_t)p.M * 3;
    } else {
        head = rn_ray_head(a.hdr + 2 * r);
        words = reinterpret_cast<const uint2 *>(a.codes + r * (int64_t)p.code_stride);
    }
    StepCount before = {0, 0, 0};
    float lsum = 0.f;
    float bestv = -INFINITY;
    int besti = 0, bestx = 0, besty = 0, bestz = 0;
    for (int c = 0; c < nch; c++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = c * RN_CHUNK + 32 * j + lane;
            int vx = 0, vy = 0, vz = 0;
            if (kAos) {
                if (i < L) { vx = __ldg(idx_row + 3 * i); vy = __ldg(idx_row + 3 * i + 1); vz = __ldg(idx_row + 3 * i + 2); }
            } else {
                const uint2 cw = __ldg(words + c * 4 + j);
                rn_decode_pair(head, cw.x, cw.y, lane, before, vx, vy, vz);
            }
            if (i < L) {
                if (!kAos && a.lin) a.lin[r * (int64_t)p.row_stride + i] = rn_brick(p, vx, vy, vz);
                float cc[3] = {__ldg(a.axes + vx), __ldg(a.axes + p.gx + vy), __ldg(a.axes + p.gx + p.gy + vz)};
                float sum = 0.f;
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    float vd = cc[q];
                    vd -= rs[q];
                    sum += ray[q] * vd;
                }
                const float tt = rn_clampf(sum / ray_norm, 1e-4f, 1 - 1e-4f);
                // stateless form of the reference's persistent two-pointer bracket: the smallest
                // left with t - (left+1)*step <= 0, searched upwards from a safe lower bound
                int left = max(0, (int)(tt * fDm1) - 2);
                while (tt - (0.0f + (float)(left + 1) * pstep) > 0 && tt - (0.0f + (float)left * pstep) > 0) left++;
                const float left_d = fabsf(tt - (0.0f + (float)left * pstep));
                const float right_d = fabsf(tt - (0.0f + (float)(left + 1) * pstep));
                // 1 - l/(l+r) = r/(l+r) and 1 - r/(l+r) = l/(l+r)
                const float inv = 1.0f / (left_d + right_d);
                const float out = (right_d * inv) * sS[left] + (left_d * inv) * sS[left + 1];
                sVal[i] = out;
                lsum += out;
                if (out > bestv) { bestv = out; besti = i; bestx = vx; besty = vy; bestz = vz; }
            }
        }
    }
    const float inv_sr = 1.0f / rn_warp_sum(lsum);
    __syncwarp();
    // normalise; clip + renormalise (mrf_np.py:4-8)
    float csum = 0.f;
    for (int i = lane; i < L; i += 32) {
        float v = sVal[i] * inv_sr;
        if (a.S_vox) a.S_vox[r * (int64_t)p.row_stride + i] = v;
        v = rn_clampf(v, 1e-5f, 0.99999f);
        sVal[i] = v;
        csum += v;
    }
    if (a.s_hat) {
        const float inv_c = 1.0f / rn_warp_sum(csum);
        __syncwarp();
        float *out_row = a.s_hat + r * (int64_t)p.row_stride;
        if (kAos) {
            for (int i = lane; i < L; i += 32) out_row[i] = sVal[i] * inv_c;
        } else {   // resident rows are 16-byte aligned and padded to whole chunks
            for (int i = 4 * lane; i < L; i += 128) {   // the last quad is zero-filled beyond L (bp2_kernel relies on it)
                float4 v = *reinterpret_cast<const float4 *>(sVal + i);
                v.x *= inv_c;
                v.y = (i + 1 < L) ? v.y * inv_c : 0.f;
                v.z = (i + 2 < L) ? v.z * inv_c : 0.f;
                v.w = (i + 3 < L) ? v.w * inv_c : 0.f;
                rn_st_stream4(out_row + i, v);
            }
        }
    }
    if (a.depth_vox) {   // mvcnn_with_ray_marching...py:280-312: first arg-max over the M slots
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            float ov = __shfl_xor_sync(RN_FULL_MASK, bestv, d);
            int oi = __shfl_xor_sync(RN_FULL_MASK, besti, d);
            int ox = __shfl_xor_sync(RN_FULL_MASK, bestx, d);
            int oy = __shfl_xor_sync(RN_FULL_MASK, besty, d);
            int oz = __shfl_xor_sync(RN_FULL_MASK, bestz, d);
            if (ov > bestv || (ov == bestv && oi < besti)) { bestv = ov; besti = oi; bestx = ox; besty = oy; bestz = oz; }
        }
        if (lane == 0) {
            float cc[3] = {__ldg(a.axes + bestx), __ldg(a.axes + p.gx + besty), __ldg(a.axes + p.gx + p.gy + bestz)};
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 3; i++) { float dd = cc[i] - sC[i]; sum += dd * dd; }
            a.depth_vox[r] = sqrtf(sum);
        }
    }
}

// =======================================================================================
// small utility kernels
// =======================================================================================
__global__ void fill_kernel(float *dst, float v, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t n4 = n >> 2;
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        for (int64_t k = i; k < n4; k += stride) d4[k] = make_float4(v, v, v, v);
        for (int64_t k = (n4 << 2) + i; k < n; k += stride) dst[k] = v;
    } else {
        for (int64_t k = i; k < n; k += stride) dst[k] = v;
    }
}

__global__ void add_prior_kernel(float *acc, float prior, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = i; k < n; k += stride) acc[k] = prior + acc[k];
}

// mrf_np.py:233-240
__global__ void occupancy_kernel(const float *acc, float *out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = i; k < n; k += stride) {
        float x = acc[k];
        float e = expf(-fabsf(x));
        out[k] = ((x >= 0.f) ? 1.0f : e) / (1.0f + e);
    }
}

__global__ void axis_centres_kernel(RnDev p, const float *voxel_grid, float *axes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t sx = (int64_t)p.gy * p.gz * 3, sy = (int64_t)p.gz * 3, sz = 3;
    if (i < p.gx) axes[i] = voxel_grid[i * sx + 0];
    else if (i < p.gx + p.gy) axes[i] = voxel_grid[(i - p.gx) * sy + 1];
    else if (i < p.gx + p.gy + p.gz) axes[i] = voxel_grid[(i - p.gx - p.gy) * sz + 2];
}

__global__ void max_count_kernel(const int32_t *count, int64_t n, int32_t *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int m = 0;
    for (int64_t k = i; k < n; k += stride) m = max(m, count[k]);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m = max(m, __shfl_xor_sync(RN_FULL_MASK, m, d));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// a1 stand-alone (thread per ray)
__global__ void sample_in_bbox_kernel(RnDev p, const int32_t *ids, const float *Pinv, const float *C, float *s,
                                      float *e, int64_t n) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float pi[12], c[3], rs[3], re[3];
#pragma unroll
    for (int i = 0; i < 12; i++) pi[i] = __ldg(Pinv + i);
#pragma unroll
    for (int i = 0; i < 3; i++) c[i] = __ldg(C + i);
    rn_sample_in_bbox(__ldg(ids + r), p, pi, c, rs, re);
#pragma unroll
    for (int i = 0; i < 3; i++) { s[3 * r + i] = rs[i]; e[3 * r + i] = re[i]; }
}

// batch_sample_points_in_bbox (sampling_schemes.cu:92-122), thread per ray
__global__ void sample_points_kernel(RnDev p, const int32_t *ids, const float *Pinv, const float *C, float *pts,
                                     int64_t n) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float pi[12], c[3], rs[3], re[3];
#pragma unroll
    for (int i = 0; i < 12; i++) pi[i] = __ldg(Pinv + i);
#pragma unroll
    for (int i = 0; i < 3; i++) c[i] = __ldg(C + i);
    rn_sample_in_bbox(__ldg(ids + r), p, pi, c, rs, re);
    float4 *out = reinterpret_cast<float4 *>(pts) + r * (int64_t)p.D;
    for (int k = 0; k < p.D; k++) {
        float4 q;
        q.x = rs[0] + (float)k * (re[0] - rs[0]) / (float)(p.D - 1);
        q.y = rs[1] + (float)k * (re[1] - rs[1]) / (float)(p.D - 1);
        q.z = rs[2] + (float)k * (re[2] - rs[2]) / (float)(p.D - 1);
        q.w = 1.0f;
        out[k] = q;
    }
}

// Stand-alone a4 on precomputed S (planes_voxels_mapping.cu:6-118): warp per ray, a lane per voxel.  The reference
// walks two persistent plane pointers along the ray; t never decreases along a ray (every DDA step moves the voxel
// centre along the ray's own direction), so they stop at the smallest right >= 1 whose float test
// t - right * step <= 0 holds -- found per voxel from ceil(t / step) with the SAME float expressions deciding.
__global__ void __launch_bounds__(128) planes_to_voxels_kernel(RnDev p, const float *axes, const int32_t *idx, const int32_t *cnt,
                                                               const float *starts, const float *ends, const float *S, float *S_new,
                                                               int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (r >= n) return;
    const int L = cnt[r];
    const int32_t *row = idx + r * (int64_t)p.M * 3;
    const float *Sr = S + r * (int64_t)p.D;
    float *out = S_new + r * (int64_t)p.M;
    float rs[3], ray[3];
    for (int i = 0; i < 3; i++) { rs[i] = starts[3 * r + i]; ray[i] = ends[3 * r + i] - rs[i]; }
    float ray_norm = 0.f;
    for (int i = 0; i < 3; i++) ray_norm += ray[i] * ray[i];
    const float step = (1.0f - 0.0f) / (float)(p.D - 1);
    float part = 0.f;
    for (int i = lane; i < L; i += 32) {
        float cc[3] = {axes[row[3 * i]], axes[p.gx + row[3 * i + 1]], axes[p.gx + p.gy + row[3 * i + 2]]};
        float sum = 0.f;
        for (int j = 0; j < 3; j++) { float vd = cc[j]; vd -= rs[j]; sum += ray[j] * vd; }
        const float t = rn_clampf(sum / ray_norm, 1e-4f, 1 - 1e-4f);
        int right = min(max((int)ceilf(t / step), 1), p.D - 1);
        while (right > 1 && t - (0.0f + (float)(right - 1) * step) <= 0) right--;
        while (right < p.D - 1 && t - (0.0f + (float)right * step) > 0) right++;
        const float left_d = fabsf(t - (0.0f + (float)(right - 1) * step)), right_d = fabsf(t - (0.0f + (float)right * step));
        const float c1 = (float)(1.0 - (double)(left_d / (left_d + right_d)));
        const float c2 = (float)(1.0 - (double)(right_d / (left_d + right_d)));
        const float v = c1 * Sr[right - 1] + c2 * Sr[right];
        out[i] = v;
        part += v;
    }
    const float srsum = rn_warp_sum(part);
    for (int i = lane; i < L; i += 32) out[i] = out[i] / srsum;   // each lane re-reads its own values
}
