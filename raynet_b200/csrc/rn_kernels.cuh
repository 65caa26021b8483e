// rn_kernels.cuh -- the sm_100a kernels of the RayNet hot path.
//
//   dda_kernel      thread-per-ray Amanatides-Woo traversal (a3), optionally fused with
//                   sample_in_bbox (a1).  Emits either the reference's int32 [M][3]
//                   lists or the resident 2-bit step codes.
//   simmap_kernel   warp-per-ray: sample_in_bbox (a1) -> plane-sweep similarity + softmax
//                   (a2) -> plane->voxel interpolation (a4) -> clip_and_renorm.
//   bp_kernel       warp-per-ray ray-potential sum-product sweep (a5, a6): forward
//                   product / prefix scans and a backward suffix scan with warp
//                   shuffles, 128-bit row loads/stores, RED scatter-add into the grid.
//   depth_kernel    warp-per-ray depth re-estimation + arg-max -> depth (a8, a9).
//
// Work decomposition for the warp-per-ray kernels: a ray of L voxels is cut into chunks
// of 128 consecutive voxels; in a chunk lane l owns voxels 4l..4l+3, so the per-ray rows
// (s_hat, messages) are read and written as one 128-bit access per lane, fully coalesced.
// Scans are "4 sequential + one 5-step warp scan", i.e. 5 shuffles per 128 voxels per
// scan.  Everything a ray needs between its forward and backward phase stays in
// registers (NCH chunks, template parameter chosen on the host from the longest ray).
#pragma once

#include "rn_common.cuh"

// =======================================================================================
// a3. DDA  (ray_tracing.pyx:99-199 == ray_tracing.cu:15-142), thread per ray
// =======================================================================================
template <bool kCodes>
__device__ __forceinline__ int rn_dda(const RnDev &p, const float *rs_in, const float *re_in,
                                      int32_t *idx_row, uint32_t *hdr, uint32_t *code_words) {
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
    if (kCodes) {
        hdr[0] = 0;
        hdr[1] = 0;
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
    uint32_t word = 3u;   // voxel 0: "no step"
    if (kCodes) {
        hdr[0] = (uint32_t)cur[0] | ((uint32_t)cur[1] << 16);
        hdr[1] = (uint32_t)cur[2] | ((step[0] < 0 ? 1u : 0u) << 16) | ((step[1] < 0 ? 1u : 0u) << 17) |
                 ((step[2] < 0 ? 1u : 0u) << 18);
    } else {
        idx_row[0] = cur[0];
        idx_row[1] = cur[1];
        idx_row[2] = cur[2];
    }
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
        if (kCodes) {
            int pos = ii & 15;
            word = (pos == 0) ? (uint32_t)a : (word | ((uint32_t)a << (2 * pos)));
            if (pos == 15) code_words[ii >> 4] = word;
        } else {
            idx_row[3 * ii + 0] = cur[0];
            idx_row[3 * ii + 1] = cur[1];
            idx_row[3 * ii + 2] = cur[2];
        }
        ii++;
    }
    if (kCodes) {
        int lastpos = (ii - 1) & 15;
        if (lastpos != 15) {   // flush the partial word, padding the tail with "no step"
            word |= (0xffffffffu << (2 * (lastpos + 1)));
            code_words[(ii - 1) >> 4] = word;
        }
    }
    return ii;
}

struct DdaArgs {
    const int32_t *ray_idxs;   // if non-null: start/end come from sample_in_bbox (and are written out)
    const float *P_inv, *centre;
    float *starts, *ends;      // inputs when ray_idxs == null; optional outputs otherwise
    int32_t *idx;              // [n][M][3]      (kCodes == false)
    uint32_t *hdr;             // [n][2]         (kCodes == true)
    uint8_t *codes;            // [n][code_stride]
    int32_t *count;            // [n]
    int64_t n_rays;
};

template <bool kCodes>
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
    int c;
    if (kCodes)
        c = rn_dda<true>(p, rs, re, nullptr, a.hdr + 2 * r, (uint32_t *)(a.codes + r * (int64_t)p.code_stride));
    else
        c = rn_dda<false>(p, rs, re, a.idx + r * (int64_t)p.M * 3, nullptr, nullptr);
    a.count[r] = c;
}

// =======================================================================================
// Per-warp decoding of a ray's voxel coordinates, chunk by chunk
// =======================================================================================
struct RayDecoder {
    int x0, y0, z0;   // first voxel
    int sx, sy, sz;   // step signs
    int bx, by, bz;   // steps taken per axis before the current chunk
};

__device__ __forceinline__ void rn_decoder_init(RayDecoder &d, const uint32_t *hdr) {
    uint32_t h0 = __ldg(hdr), h1 = __ldg(hdr + 1);
    d.x0 = h0 & 0xffff;
    d.y0 = h0 >> 16;
    d.z0 = h1 & 0xffff;
    d.sx = (h1 & (1u << 16)) ? -1 : 1;
    d.sy = (h1 & (1u << 17)) ? -1 : 1;
    d.sz = (h1 & (1u << 18)) ? -1 : 1;
    d.bx = d.by = d.bz = 0;
}

// Coordinates of the lane's 4 voxels of chunk c.  Voxels >= L get coordinates of no use
// (callers mask on the voxel index).  Must be called by the full warp, chunks in order.
template <bool kAos>
__device__ __forceinline__ void rn_decode_chunk(RayDecoder &d, const uint8_t *code_row, const int32_t *idx_row,
                                                int c, int lane, int L, int vx[4], int vy[4], int vz[4]) {
    const int i0 = c * RN_CHUNK + lane * RN_VOX_PER_LANE;
    if (kAos) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int i = i0 + j;
            bool ok = i < L;
            vx[j] = ok ? __ldg(idx_row + 3 * i + 0) : 0;
            vy[j] = ok ? __ldg(idx_row + 3 * i + 1) : 0;
            vz[j] = ok ? __ldg(idx_row + 3 * i + 2) : 0;
        }
    } else {
        uint32_t cb = (i0 < L) ? rn_ld_stream_u8(code_row + c * 32 + lane) : 0xffu;
        uint32_t run = 0, cnt[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t f = (cb >> (2 * j)) & 3u;
            run += (f == 3u) ? 0u : (1u << (10 * f));
            cnt[j] = run;
        }
        uint32_t incl = rn_warp_incl_scan_u32(run, lane);
        uint32_t excl = incl - run;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t pk = excl + cnt[j];
            vx[j] = d.x0 + d.sx * (d.bx + (int)(pk & 1023u));
            vy[j] = d.y0 + d.sy * (d.by + (int)((pk >> 10) & 1023u));
            vz[j] = d.z0 + d.sz * (d.bz + (int)(pk >> 20));
        }
        uint32_t tot = __shfl_sync(RN_FULL_MASK, incl, 31);
        d.bx += (int)(tot & 1023u);
        d.by += (int)((tot >> 10) & 1023u);
        d.bz += (int)(tot >> 20);
    }
}

// Row access: the resident layout guarantees 16-byte aligned rows (vector path); the
// reference layout (arbitrary M) uses scalar accesses.
template <bool kVec>
__device__ __forceinline__ void rn_load_row4(const float *row, int i0, int L, float v[4]) {
    if (kVec) {
        if (i0 < L) {
            float4 t = rn_ld_stream4(row + i0);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
            v[0] = v[1] = v[2] = v[3] = 0.f;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = (i0 + j < L) ? row[i0 + j] : 0.f;
    }
}
template <bool kVec>
__device__ __forceinline__ void rn_store_row4(float *row, int i0, int L, const float v[4]) {
    if (kVec) {
        // rows are padded to a multiple of 4 floats: the tail of the last quad is scratch
        if (i0 < L) rn_st_stream4(row + i0, make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (i0 + j < L) row[i0 + j] = v[j];
    }
}

// =======================================================================================
// a2 + a4. similarity + plane->voxel mapping, warp per ray
// =======================================================================================
struct SimMapArgs {
    const int32_t *ray_idxs;   // non-null: rays start from sample_in_bbox; null: starts/ends are inputs
    const float *features, *P, *P_inv, *centre;
    const int32_t *view_ids;   // optional [V]: slot of each view inside `features` (null = 0..V-1)
    const float *starts_in, *ends_in;
    float *S_planes;           // [n][D]     optional out
    float *points;             // [n][D][4]  optional out
    float *depth_planes;       // [n]        optional out: |point[argmax_k S] - C|
    // mapping stage (NCH > 0)
    const float *axes;         // [Gx+Gy+Gz] voxel-centre coordinates per axis
    const int32_t *idx;        // kAos
    const uint32_t *hdr;       // !kAos
    const uint8_t *codes;      // !kAos
    const int32_t *count;
    float *S_vox;              // [n][row_stride] optional out: normalised S_voxel_space
    float *s_hat;              // [n][row_stride] optional out: clip_and_renorm(S_voxel_space)
    float *depth_vox;          // [n] optional out: |centre(argmax voxel of S_vox) - C|
    int64_t n_rays;
};

// dynamic shared memory per CTA: [V*12 P][12 P_inv][4 C][V view slots] + per warp [D*V offsets][D S]
__host__ __device__ inline size_t rn_simmap_smem_bytes(int D, int V, int warps) {
    return sizeof(float) * (size_t)(V * 12 + 16 + V) + (size_t)warps * (sizeof(int) * (size_t)D * V + sizeof(float) * (size_t)D);
}

template <int NCH, bool kAos>
__global__ void __launch_bounds__(128) simmap_kernel(RnDev p, SimMapArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warps = blockDim.x >> 5;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *sP = reinterpret_cast<float *>(smem_raw);
    float *sPinv = sP + p.V * 12;
    float *sC = sPinv + 12;
    int *sView = reinterpret_cast<int *>(sC + 4);
    int *sOffAll = sView + p.V;
    int *sOff = sOffAll + (size_t)wid * p.D * p.V;
    float *sS = reinterpret_cast<float *>(sOffAll + (size_t)warps * p.D * p.V) + (size_t)wid * p.D;

    for (int i = threadIdx.x; i < p.V * 12; i += blockDim.x) sP[i] = __ldg(a.P + i);
    if (threadIdx.x < 12) sPinv[threadIdx.x] = a.P_inv ? __ldg(a.P_inv + threadIdx.x) : 0.f;
    if (threadIdx.x < 3) sC[threadIdx.x] = a.centre ? __ldg(a.centre + threadIdx.x) : 0.f;
    if (threadIdx.x < p.V) sView[threadIdx.x] = a.view_ids ? __ldg(a.view_ids + threadIdx.x) : (int)threadIdx.x;
    __syncthreads();

    const int64_t r = (int64_t)blockIdx.x * warps + wid;
    if (r >= a.n_rays) return;

    // ---- a1: ray start / end -------------------------------------------------------
    float rs[3], re[3];
    if (a.ray_idxs) {
        rn_sample_in_bbox(__ldg(a.ray_idxs + r), p, sPinv, sC, rs, re);
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            rs[i] = __ldg(a.starts_in + 3 * r + i);
            re[i] = __ldg(a.ends_in + 3 * r + i);
        }
    }

    // ---- a2 step A: project every (plane, view) sample, one sample per lane ----------
    const int D = p.D, V = p.V, DV = p.D * p.V;
    for (int s0 = 0; s0 < DV; s0 += 32) {
        int sidx = s0 + lane;
        if (sidx < DV) {
            int k = sidx / V, v = sidx - k * V;
            float pt[3];
#pragma unroll
            for (int i = 0; i < 3; i++) pt[i] = rs[i] + (float)k * (re[i] - rs[i]) / (float)(D - 1);
            sOff[sidx] = rn_project_offset(p, sP + v * 12, sView[v], pt);
        }
    }
    __syncwarp();

    // ---- a2 step B: S_k = sum_{i<j} <f_i, f_j> = 1/2 (|sum_v f_v|^2 - sum_v |f_v|^2) ------
    // lane = feature channel; 32 planes per block are reduced across lanes with a
    // 31-shuffle transpose-reduce so that lane l ends up with plane (kb*32 + l).
    const int DPL = (D + 31) >> 5;   // planes per lane
    float Sval[4];                   // D <= 128
#pragma unroll
    for (int kb = 0; kb < 4; kb++) {
        Sval[kb] = -INFINITY;
        if (kb < DPL) {
            float part[32];
#pragma unroll
            for (int kk = 0; kk < 32; kk++) part[kk] = 0.f;
            for (int c0 = 0; c0 < p.F; c0 += 32) {
                const int ch = c0 + lane;
                const bool chok = ch < p.F;
#pragma unroll
                for (int kk = 0; kk < 32; kk++) {
                    const int k = kb * 32 + kk;
                    if (k < D) {
                        float sum = 0.f, sq = 0.f;
                        const int *offk = sOff + k * V;
                        for (int v = 0; v < V; v++) {
                            float f = chok ? __ldg(a.features + (int64_t)offk[v] + ch) : 0.f;
                            sum += f;
                            sq = fmaf(f, f, sq);
                        }
                        part[kk] += fmaf(sum, sum, -sq);
                    }
                }
            }
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const bool upper = (lane & s) != 0;
#pragma unroll
                for (int i = 0; i < s; i++) {
                    float send = upper ? part[i] : part[i + s];
                    float keep = upper ? part[i + s] : part[i];
                    part[i] = keep + __shfl_xor_sync(RN_FULL_MASK, send, s);
                }
            }
            const int k = kb * 32 + lane;
            if (k < D) Sval[kb] = (0.5f * part[0]) / (float)p.npairs;
        }
    }
    // ---- softmax over the D planes (feature_similarities.cu:109-123) -------------------
    float mx = -INFINITY;
#pragma unroll
    for (int kb = 0; kb < 4; kb++) mx = fmaxf(mx, Sval[kb]);
    mx = rn_warp_max(mx);
    float ssum = 0.f;
#pragma unroll
    for (int kb = 0; kb < 4; kb++) {
        const int k = kb * 32 + lane;
        Sval[kb] = (kb < DPL && k < D) ? expf(Sval[kb] - mx) : 0.f;
        ssum += Sval[kb];
    }
    ssum = rn_warp_sum(ssum);
#pragma unroll
    for (int kb = 0; kb < 4; kb++) {
        const int k = kb * 32 + lane;
        if (kb < DPL && k < D) {
            float v = Sval[kb] / ssum;
            Sval[kb] = v;
            sS[k] = v;
            if (a.S_planes) a.S_planes[r * (int64_t)D + k] = v;
        }
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
    if (a.depth_planes) {   // similarities.py:213-229: first arg-max over planes
        float bv = -INFINITY;
        int bk = 0;
#pragma unroll
        for (int kb = 0; kb < 4; kb++) {
            const int k = kb * 32 + lane;
            if (kb < DPL && k < D && Sval[kb] > bv) { bv = Sval[kb]; bk = k; }
        }
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
    if constexpr (NCH > 0) {
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

        RayDecoder dec;
        const uint8_t *code_row = nullptr;
        const int32_t *idx_row = nullptr;
        if (kAos) {
            idx_row = a.idx + r * (int64_t)p.M * 3;
        } else {
            rn_decoder_init(dec, a.hdr + 2 * r);
            code_row = a.codes + r * (int64_t)p.code_stride;
        }

        float val[NCH][4];
        float lsum = 0.f;
        float bestv = -INFINITY;
        int besti = 0, bestx = 0, besty = 0, bestz = 0;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            if (c < nch) {
                int vx[4], vy[4], vz[4];
                rn_decode_chunk<kAos>(dec, code_row, idx_row, c, lane, L, vx, vy, vz);
                const int i0 = c * RN_CHUNK + lane * RN_VOX_PER_LANE;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float out = 0.f;
                    if (i0 + j < L) {
                        float cc[3] = {__ldg(a.axes + vx[j]), __ldg(a.axes + p.gx + vy[j]),
                                       __ldg(a.axes + p.gx + p.gy + vz[j])};
                        float sum = 0.f;
#pragma unroll
                        for (int t = 0; t < 3; t++) {
                            float vd = cc[t];
                            vd -= rs[t];
                            sum += ray[t] * vd;
                        }
                        float t = rn_clampf(sum / ray_norm, 1e-4f, 1 - 1e-4f);
                        // stateless form of the reference's persistent two-pointer bracket:
                        // smallest left with t - (left+1)*step <= 0
                        int left = max(0, (int)floorf(t / pstep) - 2);
                        while (t - (0.0f + (float)(left + 1) * pstep) > 0 && t - (0.0f + (float)left * pstep) > 0)
                            left++;
                        float left_d = fabsf(t - (0.0f + (float)left * pstep));
                        float right_d = fabsf(t - (0.0f + (float)(left + 1) * pstep));
                        float c1 = (float)(1.0 - (double)(left_d / (left_d + right_d)));
                        float c2 = (float)(1.0 - (double)(right_d / (left_d + right_d)));
                        out = c1 * sS[left] + c2 * sS[left + 1];
                        if (out > bestv) { bestv = out; besti = i0 + j; bestx = vx[j]; besty = vy[j]; bestz = vz[j]; }
                    }
                    val[c][j] = out;
                    lsum += out;
                }
            }
        }
        const float srsum = rn_warp_sum(lsum);
        // normalise; optionally clip + renormalise (mrf_np.py:4-8)
        float csum = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            if (c < nch) {
                const int i0 = c * RN_CHUNK + lane * RN_VOX_PER_LANE;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float v = val[c][j] / srsum;
                    val[c][j] = v;
                }
                if (a.S_vox) rn_store_row4<!kAos>(a.S_vox + r * (int64_t)p.row_stride, i0, L, val[c]);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float v = (i0 + j < L) ? rn_clampf(val[c][j], 1e-5f, 0.99999f) : 0.f;
                    val[c][j] = v;
                    csum += v;
                }
            }
        }
        if (a.s_hat) {
            csum = rn_warp_sum(csum);
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                if (c < nch) {
                    const int i0 = c * RN_CHUNK + lane * RN_VOX_PER_LANE;
#pragma unroll
                    for (int j = 0; j < 4; j++) val[c][j] = val[c][j] / csum;
                    rn_store_row4<!kAos>(a.s_hat + r * (int64_t)p.row_stride, i0, L, val[c]);
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
}

// =======================================================================================
// a5 + a6. one ray-potential BP sweep, warp per ray
// =======================================================================================
struct BpArgs {
    const float *S;            // kEngine: s_hat (already clip_and_renorm'ed); else raw S_voxel_space
    const int32_t *idx;        // !kEngine
    const uint32_t *hdr;       // kEngine
    const uint8_t *codes;      // kEngine
    const int32_t *count;
    const float *acc_in;
    float *msgs;               // in/out
    float *acc_out;
    int64_t n_rays;
};

template <int NCH, bool kEngine>
__global__ void __launch_bounds__(256) bp_kernel(RnDev p, BpArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= a.n_rays) return;
    const int L = __ldg(a.count + r);
    if (L <= 1) return;   // mrf_np.py:299-301
    const int nch = (L + RN_CHUNK - 1) / RN_CHUNK;

    const float *s_row = a.S + r * (int64_t)p.row_stride;
    float *m_row = a.msgs + r * (int64_t)p.row_stride;
    RayDecoder dec;
    const uint8_t *code_row = nullptr;
    const int32_t *idx_row = nullptr;
    if (kEngine) {
        rn_decoder_init(dec, a.hdr + 2 * r);
        code_row = a.codes + r * (int64_t)p.code_stride;
    } else {
        idx_row = a.idx + r * (int64_t)p.M * 3;
    }

    // ---- phase A: issue every load of the ray (independent of the scans) -----------------
    float sv[NCH][4], ov[NCH][4];
    int lin[NCH][4];
    float rawsum = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        if (c < nch) {
            const int i0 = c * RN_CHUNK + lane * RN_VOX_PER_LANE;
            int vx[4], vy[4], vz[4];
            rn_decode_chunk<!kEngine>(dec, code_row, idx_row, c, lane, L, vx, vy, vz);
            float mv[4];
            rn_load_row4<kEngine>(s_row, i0, L, sv[c]);
            rn_load_row4<kEngine>(m_row, i0, L, mv);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool ok = i0 + j < L;
                lin[c][j] = ok ? rn_lin(p, vx[j], vy[j], vz[j]) : 0;
                float acc = ok ? rn_ld_acc(a.acc_in + lin[c][j]) : 0.f;
                // invalid voxels: o = 0 -> factor (1-o) = 1 and a = 0: neutral in every scan
                ov[c][j] = ok ? rn_occ_to_ray(acc, mv[j]) : 0.f;
                if (!kEngine) {
                    sv[c][j] = ok ? rn_clampf(sv[c][j], 1e-5f, 0.99999f) : 0.f;
                    rawsum += sv[c][j];
                } else {
                    sv[c][j] = ok ? sv[c][j] : 0.f;
                }
            }
        }
    }
    if (!kEngine) {   // clip_and_renorm on the fly (mrf_np.py:4-8, :306)
        rawsum = rn_warp_sum(rawsum);
#pragma unroll
        for (int c = 0; c < NCH; c++)
            if (c < nch) {
#pragma unroll
                for (int j = 0; j < 4; j++) sv[c][j] = sv[c][j] / rawsum;
            }
    }

    // ---- phase B: forward scans: cp_i = prod_{k<i}(1-o_k), pre_i = sum_{j<i} a_j ----------
    float cps[NCH][4], pre[NCH][4], tot[NCH];
    float carry_cp = 1.f, carry_pre = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        tot[c] = 0.f;
        if (c < nch) {
            float lp[4];
            lp[0] = 1.f - ov[c][0];
            lp[1] = lp[0] * (1.f - ov[c][1]);
            lp[2] = lp[1] * (1.f - ov[c][2]);
            lp[3] = lp[2] * (1.f - ov[c][3]);
            float inc = rn_warp_incl_scan_mul(lp[3], lane);
            float exc = __shfl_up_sync(RN_FULL_MASK, inc, 1);
            if (lane == 0) exc = 1.f;
            const float base = carry_cp * exc;
            carry_cp = carry_cp * __shfl_sync(RN_FULL_MASK, inc, 31);
            float av[4], la[4];
            cps[c][0] = base * sv[c][0];
            cps[c][1] = (base * lp[0]) * sv[c][1];
            cps[c][2] = (base * lp[1]) * sv[c][2];
            cps[c][3] = (base * lp[2]) * sv[c][3];
#pragma unroll
            for (int j = 0; j < 4; j++) av[j] = ov[c][j] * cps[c][j];
            la[0] = av[0];
            la[1] = la[0] + av[1];
            la[2] = la[1] + av[2];
            la[3] = la[2] + av[3];
            float sinc = rn_warp_incl_scan_add(la[3], lane);
            float sexc = sinc - la[3];
            const float pbase = carry_pre + sexc;
            pre[c][0] = pbase;
            pre[c][1] = pbase + la[0];
            pre[c][2] = pbase + la[1];
            pre[c][3] = pbase + la[2];
            tot[c] = __shfl_sync(RN_FULL_MASK, sinc, 31);
            carry_pre += tot[c];
        }
    }

    // ---- phase C: backward suffix scan, messages, scatter-add ----------------------------
    float carry_suf = 0.f;
#pragma unroll
    for (int c = NCH - 1; c >= 0; c--) {
        if (c < nch) {
            const int i0 = c * RN_CHUNK + lane * RN_VOX_PER_LANE;
            float av[4], ra[4];
#pragma unroll
            for (int j = 0; j < 4; j++) av[j] = ov[c][j] * cps[c][j];
            ra[3] = av[3];
            ra[2] = av[2] + ra[3];
            ra[1] = av[1] + ra[2];
            ra[0] = av[0] + ra[1];
            float rinc = rn_warp_incl_rscan_add(ra[0], lane);
            const float sbase = carry_suf + (rinc - ra[0]);
            float suf[4] = {sbase + ra[1], sbase + ra[2], sbase + ra[3], sbase};
            float msg[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float pos = pre[c][j] + cps[c][j];
                float neg = pre[c][j] + suf[j] / (1.f - ov[c][j]);
                msg[j] = logf(pos) - logf(neg);   // == log p - log(1-p), p = pos/(pos+neg)
            }
            rn_store_row4<kEngine>(m_row, i0, L, msg);
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (i0 + j < L) rn_red_add(a.acc_out + lin[c][j], msg[j]);
            carry_suf += tot[c];
        }
    }
}

// =======================================================================================
// a8 + a9. depth re-estimation (+ arg-max -> depth), warp per ray
// =======================================================================================
struct DepthArgs {
    const float *S;            // kEngine: s_hat; else raw S_voxel_space
    const int32_t *idx;
    const uint32_t *hdr;
    const uint8_t *codes;
    const int32_t *count;
    const float *acc;
    const float *msgs;
    const float *axes;         // needed when depth_map != null
    const float *centre;
    float *S_new;              // [n][row_stride] optional out (normalised, zero beyond count)
    float *depth_map;          // [n] optional out
    int64_t n_rays;
};

template <int NCH, bool kEngine>
__global__ void __launch_bounds__(256) depth_kernel(RnDev p, DepthArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= a.n_rays) return;
    const int L = __ldg(a.count + r);
    const int nch = (max(L, 0) + RN_CHUNK - 1) / RN_CHUNK;
    const float *s_row = a.S + r * (int64_t)p.row_stride;
    const float *m_row = a.msgs + r * (int64_t)p.row_stride;
    RayDecoder dec;
    dec.x0 = dec.y0 = dec.z0 = 0;
    const uint8_t *code_row = nullptr;
    const int32_t *idx_row = nullptr;
    if (kEngine) {
        rn_decoder_init(dec, a.hdr + 2 * r);
        code_row = a.codes + r * (int64_t)p.code_stride;
    } else {
        idx_row = a.idx + r * (int64_t)p.M * 3;
    }
    int fx = 0, fy = 0, fz = 0;   // voxel of slot 0 (what the reference reads for an all-zero row)
    if (L >= 1) {
        if (kEngine) { fx = dec.x0; fy = dec.y0; fz = dec.z0; }
        else { fx = __ldg(idx_row); fy = __ldg(idx_row + 1); fz = __ldg(idx_row + 2); }
    }

    float av[NCH][4];
    float bestv = -INFINITY;
    int besti = 0, bestx = fx, besty = fy, bestz = fz;
    float asum = 0.f;
    if (L > 1) {   // mrf_np.py:376-377: rays with count <= 1 keep an all-zero row
        float sv[NCH][4], ov[NCH][4];
        int cx[NCH][4], cy[NCH][4], cz[NCH][4];
        float rawsum = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            if (c < nch) {
                const int i0 = c * RN_CHUNK + lane * RN_VOX_PER_LANE;
                rn_decode_chunk<!kEngine>(dec, code_row, idx_row, c, lane, L, cx[c], cy[c], cz[c]);
                float mv[4];
                rn_load_row4<kEngine>(s_row, i0, L, sv[c]);
                rn_load_row4<kEngine>(m_row, i0, L, mv);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const bool ok = i0 + j < L;
                    float acc = ok ? rn_ld_acc(a.acc + rn_lin(p, cx[c][j], cy[c][j], cz[c][j])) : 0.f;
                    ov[c][j] = ok ? rn_occ_to_ray(acc, mv[j]) : 0.f;
                    if (!kEngine) {
                        sv[c][j] = ok ? rn_clampf(sv[c][j], 1e-5f, 0.99999f) : 0.f;
                        rawsum += sv[c][j];
                    } else {
                        sv[c][j] = ok ? sv[c][j] : 0.f;
                    }
                }
            }
        }
        if (!kEngine) {
            rawsum = rn_warp_sum(rawsum);
#pragma unroll
            for (int c = 0; c < NCH; c++)
                if (c < nch) {
#pragma unroll
                    for (int j = 0; j < 4; j++) sv[c][j] = sv[c][j] / rawsum;
                }
        }
        float carry_cp = 1.f;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            if (c < nch) {
                const int i0 = c * RN_CHUNK + lane * RN_VOX_PER_LANE;
                float lp[4];
                lp[0] = 1.f - ov[c][0];
                lp[1] = lp[0] * (1.f - ov[c][1]);
                lp[2] = lp[1] * (1.f - ov[c][2]);
                lp[3] = lp[2] * (1.f - ov[c][3]);
                float inc = rn_warp_incl_scan_mul(lp[3], lane);
                float exc = __shfl_up_sync(RN_FULL_MASK, inc, 1);
                if (lane == 0) exc = 1.f;
                const float base = carry_cp * exc;
                carry_cp = carry_cp * __shfl_sync(RN_FULL_MASK, inc, 31);
                av[c][0] = ov[c][0] * (base * sv[c][0]);
                av[c][1] = ov[c][1] * ((base * lp[0]) * sv[c][1]);
                av[c][2] = ov[c][2] * ((base * lp[1]) * sv[c][2]);
                av[c][3] = ov[c][3] * ((base * lp[2]) * sv[c][3]);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (i0 + j < L) {
                        asum += av[c][j];
                        if (av[c][j] > bestv) {
                            bestv = av[c][j]; besti = i0 + j;
                            bestx = cx[c][j]; besty = cy[c][j]; bestz = cz[c][j];
                        }
                    }
                }
            }
        }
        asum = rn_warp_sum(asum);
    }
    if (a.S_new) {   // normalised distribution, zero beyond count (mrf_np.py:370, :378)
        float *o_row = a.S_new + r * (int64_t)p.row_stride;
        if (L > 1) {
#pragma unroll
            for (int c = 0; c < NCH; c++)
                if (c < nch) {
                    const int i0 = c * RN_CHUNK + lane * RN_VOX_PER_LANE;
                    float v[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) v[j] = av[c][j] / asum;
                    rn_store_row4<kEngine>(o_row, i0, L, v);
                }
        }
        const int zfrom = (L > 1) ? L : 0;
        if (!kEngine)
            for (int i = zfrom + lane; i < p.M; i += 32) o_row[i] = 0.f;
    }
    if (a.depth_map) {   // raynet_fp.py:193-226
        if (L > 1) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                float ovv = __shfl_xor_sync(RN_FULL_MASK, bestv, d);
                int oi = __shfl_xor_sync(RN_FULL_MASK, besti, d);
                int ox = __shfl_xor_sync(RN_FULL_MASK, bestx, d);
                int oy = __shfl_xor_sync(RN_FULL_MASK, besty, d);
                int oz = __shfl_xor_sync(RN_FULL_MASK, bestz, d);
                if (ovv > bestv || (ovv == bestv && oi < besti)) { bestv = ovv; besti = oi; bestx = ox; besty = oy; bestz = oz; }
            }
        }
        if (lane == 0) {
            float cc[3] = {__ldg(a.axes + bestx), __ldg(a.axes + p.gx + besty), __ldg(a.axes + p.gx + p.gy + bestz)};
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 3; i++) { float dd = cc[i] - __ldg(a.centre + i); sum += dd * dd; }
            a.depth_map[r] = sqrtf(sum);
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

// Expand resident step codes into the reference's dense int32 [M][3] lists (thread per ray).
__global__ void expand_indices_kernel(RnDev p, const uint32_t *hdr, const uint8_t *codes, const int32_t *count,
                                      int32_t *idx, int64_t n_rays) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    int L = count[r];
    int32_t *row = idx + r * (int64_t)p.M * 3;
    uint32_t h0 = hdr[2 * r], h1 = hdr[2 * r + 1];
    int x = h0 & 0xffff, y = h0 >> 16, z = h1 & 0xffff;
    int sx = (h1 & (1u << 16)) ? -1 : 1, sy = (h1 & (1u << 17)) ? -1 : 1, sz = (h1 & (1u << 18)) ? -1 : 1;
    const uint8_t *crow = codes + r * (int64_t)p.code_stride;
    for (int i = 0; i < p.M; i++) {
        if (i < L) {
            uint32_t f = (crow[i >> 2] >> (2 * (i & 3))) & 3u;
            if (f == 0) x += sx;
            else if (f == 1) y += sy;
            else if (f == 2) z += sz;
            row[3 * i] = x; row[3 * i + 1] = y; row[3 * i + 2] = z;
        } else {
            row[3 * i] = 0; row[3 * i + 1] = 0; row[3 * i + 2] = 0;
        }
    }
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

// Stand-alone a4 on precomputed S (planes_voxels_mapping.cu:6-118): thread per ray,
// sequential over the ray with the reference's persistent two-pointer bracket.
__global__ void planes_to_voxels_kernel(RnDev p, const float *axes, const int32_t *idx, const int32_t *cnt,
                                        const float *starts, const float *ends, const float *S, float *S_new,
                                        int64_t n) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
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
    int left = 0, right = 1;
    float srsum = 0.f;
    for (int i = 0; i < L; i++) {
        float cc[3] = {axes[row[3 * i]], axes[p.gx + row[3 * i + 1]], axes[p.gx + p.gy + row[3 * i + 2]]};
        float sum = 0.f;
        for (int j = 0; j < 3; j++) { float vd = cc[j]; vd -= rs[j]; sum += ray[j] * vd; }
        float t = rn_clampf(sum / ray_norm, 1e-4f, 1 - 1e-4f);
        float left_d = t - (0.0f + (float)left * step), right_d = t - (0.0f + (float)right * step);
        while (left_d > 0 && right_d > 0) {
            left++; right++;
            left_d = t - (0.0f + (float)left * step);
            right_d = t - (0.0f + (float)right * step);
        }
        left_d = fabsf(left_d); right_d = fabsf(right_d);
        float c1 = (float)(1.0 - (double)(left_d / (left_d + right_d)));
        float c2 = (float)(1.0 - (double)(right_d / (left_d + right_d)));
        float v = c1 * Sr[left] + c2 * Sr[right];
        out[i] = v;
        srsum += v;
    }
    for (int i = 0; i < L; i++) out[i] = out[i] / srsum;
}
