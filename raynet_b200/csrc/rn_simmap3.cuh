// rn_simmap3.cuh -- front end of the resident pipeline, third version, F = 32, one warp per ray, as TWO
// kernels: simscore3_kernel = plane-sweep similarity + softmax (a2) -> S_planes (256 B per ray);
// planemap3_kernel = plane->voxel interpolation (a4) + clip_and_renorm -> s_hat, lin.  Split because
// the similarity is bound by L2->SM gather bandwidth and wants every warp in the gather loop, while
// the mapping is issue-bound and needs 3 KB of shared memory per ray.
//
// simmap_kernel (rn_kernels.cuh, profiles/r01_simmap_source_lines.txt) is instruction-bound: 8300
// warp instructions per ray at 76 % issue utilisation, a third of them the 576 (plane, view)
// projections with three IEEE divisions for the plane point and two for the pixel, each
// repeated per view.  Here
//   * a lane owns PLANES (k = lane, lane + 32, ...): the plane point is formed once per plane
//     (the oracle's IEEE arithmetic) and then projected into the V views;
//   * the pixel division runs on the fast path q = o * rcp.approx(nz): q is within 2 ulp of the
//     true quotient, so whenever q is further than that from every half-integer it rounds to
//     the same pixel as the oracle's IEEE quotient; the few samples that are too close (or not
//     finite) take the exact path.  Integer decisions stay bit-exact, ~10x fewer instructions;
//   * the feature gathers accumulate with packed FADD2 / FFMA2 (two channels per instruction), are
//     addressed with 32-bit byte offsets on a per-lane 64-bit base, and are unrolled per view count
//     (template VT) so that all V loads of a plane group are in flight together; the reference
//     view, whose 64 samples all land on the ray's own pixel, is gathered once;
//   * per-axis tables of voxel-centre coordinates and of bricked accumulator offsets live in
//     shared memory (one CTA serves RN_SM3_RAYS_PER_WARP rays per warp), the plane bracket of
//     planes_voxels_mapping.cu is evaluated in closed form (the interpolant is continuous in t).
#pragma once

#include "rn_kernels.cuh"

#ifndef RN_SIMSCORE_WIDE
#define RN_SIMSCORE_WIDE 16      // warps (rays) of the wide CTA flavour: 16 (4x4 pixel patch) or 32 (8x4); 0 = never
#endif
#ifndef RN_SIMSCORE_SYNC
#define RN_SIMSCORE_SYNC 0       // > 0: the CTA's warps re-align every this many plane groups (bar.sync)
#endif
#ifndef RN_SIMSCORE_UNROLL
#define RN_SIMSCORE_UNROLL 2     // plane groups (of 4 planes x V gathers) in flight per warp
#endif
#ifndef RN_SM3_RAYS_PER_WARP
#define RN_SM3_RAYS_PER_WARP 4
#endif

__device__ __forceinline__ uint64_t rn_pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void rn_unpack2(uint64_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t rn_add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t rn_fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// dynamic shared memory (floats)
//   simscore3_kernel: per CTA [V*12 P][V view bases], per warp [D*V feature offsets][D + 4 plane scores]
//   planemap3_kernel: per CTA [3 G axes][3 G brick offsets], per warp [D + 4 plane scores][val_stride voxel values]
__host__ __device__ inline size_t rn_simscore3_cta_words(int V) { return (size_t)V * 12 + (size_t)((V + 3) & ~3); }
__host__ __device__ inline size_t rn_simscore3_warp_words(int D, int V) { return (size_t)D * ((V + 3) & ~3) + (size_t)((D + 7) & ~3); }
__host__ __device__ inline size_t rn_planemap3_cta_words(int gsum) { return 2 * (size_t)((gsum + 1) & ~1); }
__host__ __device__ inline size_t rn_planemap3_warp_words(int D, int val_stride) { return 2 * (size_t)((D + 3) & ~1) + (size_t)val_stride; }

// feature_similarities.cu:42-61 on an already rounded pixel
__device__ __forceinline__ int rn_feature_offset(const RnDev &p, int base, int fx, int fy) {
    fx = min(max(fx, 0), p.W);
    fy = min(max(fy, 0), p.H);
    if (fx == 0 || fy == 0) fx = fy = 0;
    return base + (fy * p.fw + fx) * p.F;
}

// Step 2 of the similarity: 8 lanes cover one 32-channel vector (LDG.128 each), 4 planes per warp
// instruction.  sOff holds BYTE offsets (uint32) into the feature volume, so an address is the lane's
// 64-bit base plus a 32-bit offset: two integer instructions instead of the four of a sign-extended
// element index.  kRef: the reference view's vector (the same for every plane) was fetched once and
// seeds the accumulators; separate instantiations keep predicated copies out of the loop.
template <int VT, bool kRef>
__device__ __forceinline__ void rn_plane_scores(const RnDev &p, const SimMapArgs &a, const int *sOff, float *sS,
                                                int ref_off, int lane, const int D /* planes of this pass */) {
    const int V = VT ? VT : p.V, VP = (V + 3) & ~3;
    const float inv_pairs = 0.5f / (float)p.npairs;
    const int g = lane >> 3;
    const char *featb = reinterpret_cast<const char *>(a.features) + (lane & 7) * 16;
    auto vec = [&](int off) { return __ldg(reinterpret_cast<const float4 *>(featb + (uint32_t)off)); };
    uint64_t r01 = 0, r23 = 0, rq01 = 0, rq23 = 0;
    if (kRef) {
        const float4 f = vec(ref_off);
        r01 = rn_pack2(f.x, f.y); r23 = rn_pack2(f.z, f.w);
        rq01 = rn_fma2(r01, r01, 0); rq23 = rn_fma2(r23, r23, 0);
    }
    constexpr int V0 = kRef ? 1 : 0;
    constexpr int kUnroll = RN_SIMSCORE_UNROLL;
#pragma unroll kUnroll
    for (int k0 = 0; k0 < D; k0 += 4) {
        if (RN_SIMSCORE_SYNC > 0 && (k0 % (4 * RN_SIMSCORE_SYNC)) == 0) __syncthreads();
        const int k = min(k0 + g, D - 1);
        uint64_t s01 = r01, s23 = r23, q01 = rq01, q23 = rq23;
        if constexpr (VT > 0) {
            constexpr int NG = (VT + 3) / 4;
            const int4 *offk = reinterpret_cast<const int4 *>(sOff + k * VP);
            int off[NG * 4];
#pragma unroll
            for (int gq = 0; gq < NG; gq++) {
                const int4 o4 = offk[gq];
                off[4 * gq] = o4.x; off[4 * gq + 1] = o4.y; off[4 * gq + 2] = o4.z; off[4 * gq + 3] = o4.w;
            }
            float4 f[VT];
#pragma unroll
            for (int v = V0; v < VT; v++) f[v] = vec(off[v]);
#pragma unroll
            for (int v = V0; v < VT; v++) {
                const uint64_t f01 = rn_pack2(f[v].x, f[v].y), f23 = rn_pack2(f[v].z, f[v].w);
                s01 = rn_add2(s01, f01); s23 = rn_add2(s23, f23);
                q01 = rn_fma2(f01, f01, q01); q23 = rn_fma2(f23, f23, q23);
            }
        } else {
            const int *offs = sOff + k * VP;
#pragma unroll 4
            for (int v = V0; v < V; v++) {
                const float4 fv = vec(offs[v]);
                const uint64_t f01 = rn_pack2(fv.x, fv.y), f23 = rn_pack2(fv.z, fv.w);
                s01 = rn_add2(s01, f01); s23 = rn_add2(s23, f23);
                q01 = rn_fma2(f01, f01, q01); q23 = rn_fma2(f23, f23, q23);
            }
        }
        float sx, sy, sz, sw, qx, qy, qz, qw;
        rn_unpack2(s01, sx, sy); rn_unpack2(s23, sz, sw);
        rn_unpack2(q01, qx, qy); rn_unpack2(q23, qz, qw);
        float val = fmaf(sx, sx, fmaf(sy, sy, fmaf(sz, sz, fmaf(sw, sw, -((qx + qy) + (qz + qw))))));
        val += __shfl_xor_sync(RN_FULL_MASK, val, 1);
        val += __shfl_xor_sync(RN_FULL_MASK, val, 2);
        val += __shfl_xor_sync(RN_FULL_MASK, val, 4);
        if ((lane & 7) == 0 && k0 + g < D) sS[k0 + g] = val * inv_pairs;
    }
}

// a2: plane-sweep similarity + softmax -> S_planes [n][D].  All warps of all CTAs spend their time
// in the projection / gather loops (nothing else competes for registers and shared memory), which
// is what keeps enough 128-byte feature gathers in flight to load the L2.
// Which ray of its 64-pixel tile warp `wid` of CTA `cta` serves.  With 4 warps a CTA takes 4 vertical neighbours;
// with 16 / 32 warps a compact 4x4 / 8x4 patch (x, y), so that the rays whose plane samples fall on the same pixels
// of the other views are resident on ONE SM at the same time and meet in its L1 (scratch/sim_l1b.py: modelled L1
// hit rate 0.15 -> 0.39 / 0.46 on the C3 rig).
template <int kW>
__device__ __forceinline__ int64_t rn_simscore_slot(int64_t cta, int wid, int64_t tile_len) {
    const int64_t t = cta * kW + wid;
    if (tile_len <= 0 || kW == 4) return t;
    const int in = (int)(t & 63);
    int xo, yo;
    if (kW == 16) {
        const int q = in >> 4, w = in & 15;
        xo = (q >> 1) * 4 + (w >> 2);
        yo = (q & 1) * 4 + (w & 3);
    } else {
        const int h = in >> 5, w = in & 31;
        xo = w >> 2;
        yo = h * 4 + (w & 3);
    }
    return (t & ~(int64_t)63) | (xo << 3) | yo;
}

// Warps per SM the register budget is set for: 32 (<= 64 registers) up to 11 views; beyond that the V gathers of two
// plane groups no longer fit in 64 registers (15 views: spills, loads serialised) -- 16 warps at <= 128 registers.
#ifndef RN_SIMSCORE_MANY_VIEWS
#define RN_SIMSCORE_MANY_VIEWS 12
#endif
__host__ __device__ constexpr int rn_simscore_warps_per_sm(int vt) { return (vt >= RN_SIMSCORE_MANY_VIEWS) ? 16 : 32; }

template <int VT, int kW>
__global__ void __launch_bounds__(32 * kW, rn_simscore_warps_per_sm(VT) / kW) simscore3_kernel(RnDev p, SimMapArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.D, V = VT ? VT : p.V;   // VT > 0: compile-time view count, loops fully unrolled
    float *sP = reinterpret_cast<float *>(smem_raw);
    int *sBase = reinterpret_cast<int *>(sP + V * 12);
    float *warp0 = reinterpret_cast<float *>(sBase + ((V + 3) & ~3));
    int *sOff = reinterpret_cast<int *>(warp0 + (size_t)wid * rn_simscore3_warp_words(D, V));
    float *sS = reinterpret_cast<float *>(sOff + D * ((V + 3) & ~3));

    for (int i = threadIdx.x; i < V * 12; i += blockDim.x) sP[i] = __ldg(a.P + i);
    if (threadIdx.x < V) {
        const int slot = a.view_ids ? __ldg(a.view_ids + threadIdx.x) : (int)threadIdx.x;
        sBase[threadIdx.x] = slot * p.fh * p.fw * p.F;
    }
    __syncthreads();

    const float fDm1 = (float)(D - 1);
    const int fshift = p.shift;
    // planes of this pass (all of them unless the launcher sweeps the planes in blocks, see launch_plane_scores)
    const int k_lo = a.k_hi > 0 ? a.k_lo : 0, k_hi = a.k_hi > 0 ? a.k_hi : D;
    const int Dl = k_hi - k_lo;

    const int64_t t = rn_simscore_slot<kW>(blockIdx.x, wid, a.tile_len);
    if (!RN_SIMSCORE_SYNC && t >= a.n_rays) return;
    const bool live = t < a.n_rays;
    const int64_t r = live ? rn_tiled_position(t, a.tile_len, p.H, a.tile_mode & 0xff) : 0;
    float rs[3], re[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        rs[i] = __ldg(a.starts_in + 3 * r + i);
        re[i] = __ldg(a.ends_in + 3 * r + i);
    }

    // ---- step 1: a lane owns planes k = lane, lane + 32 (two per pass); project into every view ----
    // sOff rows are padded to VP = 4 * ceil(V / 4) ints so that step 2 reads four offsets per LDS.128
    const int VP = (V + 3) & ~3;
    bool ref_same = true;   // all samples of view 0 land on one pixel (the reference view)
    int ref_off = 0;
    for (int kb = k_lo; kb < k_hi; kb += 64) {
        const int kk[2] = {kb + lane, kb + 32 + lane};
        float pt[2][3];
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int i = 0; i < 3; i++) pt[h][i] = rs[i] + (float)kk[h] * (re[i] - rs[i]) / fDm1;   // feature_similarities.cu:80-84
#pragma unroll
        for (int v = 0; v < V; v++) {
            const float4 P0 = *reinterpret_cast<const float4 *>(sP + v * 12);
            const float4 P1 = *reinterpret_cast<const float4 *>(sP + v * 12 + 4);
            const float4 P2 = *reinterpret_cast<const float4 *>(sP + v * 12 + 8);
            const int vbase = sBase[v];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                // feature_similarities.cu:10-32, operation for operation (-fmad=false: no contraction)
                float o0 = P0.x * pt[h][0]; o0 += P0.y * pt[h][1]; o0 += P0.z * pt[h][2]; o0 += P0.w * 1;
                float o1 = P1.x * pt[h][0]; o1 += P1.y * pt[h][1]; o1 += P1.z * pt[h][2]; o1 += P1.w * 1;
                float nz = P2.x * pt[h][0]; nz += P2.y * pt[h][1]; nz += P2.z * pt[h][2]; nz += P2.w * 1;
                const float rc = rn_rcp(nz);
                const float q0 = o0 * rc, q1 = o1 * rc;
                const float n0 = rintf(q0), n1 = rintf(q1);
                // safe: both quotients are further than their own error bound (2 ulp; 1e-6 relative is
                // generous) from a rounding boundary; NaN, infinities and |q| >= 5e5 fail the test
                const float e0 = fabsf(q0 - n0) + fabsf(q0) * 1e-6f, e1 = fabsf(q1 - n1) + fabsf(q1) * 1e-6f;
                int fx, fy;
                if (fmaxf(e0, e1) < 0.49999f) {
                    fx = (int)n0 + fshift;
                    fy = (int)n1 + fshift;
                } else {   // the oracle's arithmetic: IEEE quotient, round half away from zero
                    fx = (int)(roundf(o0 / nz) + (float)fshift);
                    fy = (int)(roundf(o1 / nz) + (float)fshift);
                }
                const int off = (int)((uint32_t)rn_feature_offset(p, vbase, fx, fy) * 4u);   // BYTE offset (feature volume < 4 GiB)
                if (kk[h] < k_hi) {
                    sOff[(kk[h] - k_lo) * VP + v] = off;   // (16-byte stores of a plane's offsets, conflict-free, measured slower: 12.55 vs 11.94 ms)
                    if (v == 0) {
                        if (kk[h] == k_lo + lane) ref_off = off;
                        else if (off != ref_off) ref_same = false;
                    }
                }
            }
        }
    }
    {
        const int off0 = __shfl_sync(RN_FULL_MASK, ref_off, 0);
        ref_same = __all_sync(RN_FULL_MASK, ref_same && (lane >= Dl || ref_off == off0));
        ref_off = off0;
    }
    __syncwarp();

    // ---- step 2: plane scores S_k = 1/2 (|sum_v f_v|^2 - sum_v |f_v|^2) / pairs ----------------------
    if (ref_same) rn_plane_scores<VT, true>(p, a, sOff, sS, ref_off, lane, Dl);
    else rn_plane_scores<VT, false>(p, a, sOff, sS, ref_off, lane, Dl);
    __syncwarp();
    if (a.raw_scores) {   // plane-blocked sweep: the softmax follows once every block is in (softmax_planes_kernel)
        if (live)
            for (int k = lane; k < Dl; k += 32) a.S_planes[r * (int64_t)D + k_lo + k] = sS[k];
        return;
    }

    // ---- step 3: softmax over the D planes (feature_similarities.cu:109-123) --------------------------
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
    if (live)
        for (int k = lane; k < D; k += 32) a.S_planes[r * (int64_t)D + k] = sS[k] / ssum;
}

// Softmax over the D plane scores of every ray, in place (feature_similarities.cu:109-123) -- step 3 of
// simscore3_kernel as its own kernel for the plane-blocked sweep; the same lane assignment and operation order, so
// the result is bit-identical to the single-pass kernel's.
__global__ void __launch_bounds__(128) softmax_planes_kernel(float *S_planes, int64_t n_rays, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (r >= n_rays) return;
    float *row = S_planes + r * (int64_t)D;
    float mx = -INFINITY;
    for (int k = lane; k < D; k += 32) mx = fmaxf(mx, row[k]);
    mx = rn_warp_max(mx);
    float ssum = 0.f;
    float ev[4];   // D <= 128
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int k = lane + 32 * j;
        ev[j] = (k < D) ? expf(row[k] - mx) : 0.f;
        if (k < D) ssum += ev[j];
    }
    ssum = rn_warp_sum(ssum);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int k = lane + 32 * j;
        if (k < D) row[k] = ev[j] / ssum;
    }
}

// a4: plane -> voxel mapping (planes_voxels_mapping.cu:6-92) + clip_and_renorm (mrf_np.py:4-8) from
// S_planes -> s_hat, lin.  One warp per ray, RN_SM3_RAYS_PER_WARP rays per warp so that the per-axis
// tables in shared memory are amortised.
__global__ void __launch_bounds__(128) planemap3_kernel(RnDev p, SimMapArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.D;
    const int gsum = p.gx + p.gy + p.gz;
    int2 *sTab = reinterpret_cast<int2 *>(smem_raw);   // per axis entry: (bricked offset contribution, centre coordinate)
    float *warp0 = reinterpret_cast<float *>(sTab + ((gsum + 1) & ~1));
    float2 *sS2 = reinterpret_cast<float2 *>(warp0 + (size_t)wid * rn_planemap3_warp_words(D, a.val_stride));   // (S_k, S_k+1 - S_k)
    float *sVal = reinterpret_cast<float *>(sS2 + ((D + 3) & ~1));
    for (int i = threadIdx.x; i < gsum; i += blockDim.x) {
        const int b = (i < p.gx) ? rn_brick_fx(p, i) : (i < p.gx + p.gy) ? rn_brick_fy(p, i - p.gx) : rn_brick_fz(i - p.gx - p.gy);
        sTab[i] = make_int2(b, __float_as_int(__ldg(a.axes + i)));
    }
    __syncthreads();
    const float fDm1 = (float)(D - 1);
    const float pstep = (1.0f - 0.0f) / fDm1;

    for (int it = 0; it < RN_SM3_RAYS_PER_WARP; it++) {
        const int64_t r = ((int64_t)blockIdx.x * RN_SM3_RAYS_PER_WARP + it) * 4 + wid;   // plain ray order: rows stream
        if (r >= a.n_rays) break;
        const int L = __ldg(a.count + r);
        if (L <= 0) continue;
        float rs[3], re[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            rs[i] = __ldg(a.starts_in + 3 * r + i);
            re[i] = __ldg(a.ends_in + 3 * r + i);
        }
        __syncwarp();   // the previous ray's reads of this warp's buffers are done
        for (int k = lane; k < D; k += 32) {
            const float s0 = __ldg(a.S_planes + r * (int64_t)D + k);
            const float s1 = (k + 1 < D) ? __ldg(a.S_planes + r * (int64_t)D + k + 1) : s0;
            sS2[k] = make_float2(s0, s1 - s0);
        }
        __syncwarp();
        const int nch = (L + RN_CHUNK - 1) / RN_CHUNK;
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
        for (int c = 0; c < nch; c++) {
            const uint4 cwA = __ldg(reinterpret_cast<const uint4 *>(words + c * 4));
            const uint4 cwB = __ldg(reinterpret_cast<const uint4 *>(words + c * 4 + 2));
            const uint32_t lo[4] = {cwA.x, cwA.z, cwB.x, cwB.z}, hi[4] = {cwA.y, cwA.w, cwB.y, cwB.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = c * RN_CHUNK + 32 * j + lane;
                int vx, vy, vz;
                rn_decode_pair(head, lo[j], hi[j], lane, before, vx, vy, vz);
                if (i < L) {
                    const int2 ex = sTab[vx], ey = sTab[p.gx + vy], ez = sTab[p.gx + p.gy + vz];   // (brick offset, centre)
                    lin_row[i] = ex.x + ey.x + ez.x;
                    // t = <centre - start, ray> / |ray|^2 (planes_voxels_mapping.cu:36-52) with the division
                    // folded into the ray (tolerance-gated value: the interpolation is continuous in t)
                    const float sum = (__int_as_float(ex.y) - rs[0]) * rayn[0] + (__int_as_float(ey.y) - rs[1]) * rayn[1] +
                                      (__int_as_float(ez.y) - rs[2]) * rayn[2];
                    const float tt = rn_clampf(sum, 1e-4f, 1 - 1e-4f);
                    // the reference's persistent two-pointer bracket (planes_voxels_mapping.cu:54-70) picks the
                    // planes left, left + 1 around t; the interpolant is continuous and piecewise linear in t, so
                    // left = floor(t (D - 1)) gives the same value (to rounding) even where the two differ at a
                    // plane position.  t <= 1 - 1e-4 keeps left <= D - 2.
                    const int left = min((int)(tt * fDm1), D - 2);
                    const float2 sd = sS2[left];
                    const float out = fmaf((tt - (float)left * pstep) * fDm1, sd.y, sd.x);
                    sVal[i] = out;
                    lsum += out;
                }
            }
        }
        const float inv_sr = 1.0f / rn_warp_sum(lsum);
        __syncwarp();
        // normalise; clip + renormalise (mrf_np.py:4-8); the last quad is zero-filled beyond L
        float csum = 0.f;
        for (int i = 4 * lane; i < L; i += 128) {
            float4 v = *reinterpret_cast<const float4 *>(sVal + i);
            v.x = rn_clampf(v.x * inv_sr, 1e-5f, 0.99999f);
            v.y = (i + 1 < L) ? rn_clampf(v.y * inv_sr, 1e-5f, 0.99999f) : 0.f;
            v.z = (i + 2 < L) ? rn_clampf(v.z * inv_sr, 1e-5f, 0.99999f) : 0.f;
            v.w = (i + 3 < L) ? rn_clampf(v.w * inv_sr, 1e-5f, 0.99999f) : 0.f;
            *reinterpret_cast<float4 *>(sVal + i) = v;
            csum += (v.x + v.y) + (v.z + v.w);
        }
        const float inv_c = 1.0f / rn_warp_sum(csum);
        float *out_row = a.s_hat + r * (int64_t)p.row_stride;
        for (int i = 4 * lane; i < L; i += 128) {   // each lane re-reads its own quads
            float4 v = *reinterpret_cast<const float4 *>(sVal + i);
            v.x *= inv_c; v.y *= inv_c; v.z *= inv_c; v.w *= inv_c;
            rn_st_stream4(out_row + i, v);
        }
    }
}
