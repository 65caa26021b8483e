// rn_engine.cuh -- the resident (B200-native) ray-potential pipeline.
//
// Per-ray state lives in HBM between sweeps (include/raynet_b200.h, "Resident pipeline"):
//   hdr    uint32 [n][2]            first voxel + step signs
//   codes  2 bits / traversed voxel as BIT PLANES: for every 32 consecutive voxels one
//          (lo, hi) pair of 32-bit words; code = hi<<1 | lo: 0/1/2 = the step along x/y/z
//          that led INTO the voxel, 3 = no step (voxel 0 and padding).  Any lane can
//          turn the words into its voxel's coordinates with three popc's -- no shuffles.
//   s_hat  float32 [n][M]           clip_and_renorm(S_voxel_space)
//   msgs   float32 [n][M]
// and the two occupancy accumulators are stored BRICKED: a 128-byte line holds a 4x4x2
// (x, y, z) block of voxels made of four 32-byte sectors of 2x2x2 voxels, so a ray that
// crosses the grid in any direction touches ~3x fewer sectors than in the row-major
// layout of the reference (where only z-steps stay inside a sector).
//
// Kernels:
//   dda_codes_kernel   thread per ray: sample_in_bbox (a1) + Amanatides-Woo (a3) -> hdr, codes, count
//   bin_*_kernel       rays -> length classes (number of 128-voxel chunks), 2-D tiled order
//   bp2_kernel         warp per ray, one BP sweep (a5 + a6): the ray's rows are staged into
//                      shared memory with TMA bulk copies (cp.async.bulk + mbarrier), the
//                      accumulator is gathered lane-consecutively (sector sharing), the
//                      forward/backward scans run on 4 consecutive voxels per lane with warp
//                      shuffles, new messages leave through a TMA bulk store and fire-and-forget
//                      RED.ADD into the new accumulator.
//   depth2_kernel      warp per ray depth re-estimation + arg-max -> depth (a8 + a9)
#pragma once

#include "rn_common.cuh"

#define RN_NCLASS (RN_MAX_NCH + 1)   // class c = ceil(count / 128) for count >= 2 (1..RN_MAX_NCH); class 0 = rays BP skips

// ---- brick layout -----------------------------------------------------------------------
__host__ __device__ __forceinline__ int rn_brick_fx(const RnDev &p, int x) {
    return (x >> 2) * p.bsx + ((x >> 1) & 1) * 16 + (x & 1) * 4;
}
__host__ __device__ __forceinline__ int rn_brick_fy(const RnDev &p, int y) {
    return (y >> 2) * p.bsy + ((y >> 1) & 1) * 8 + (y & 1) * 2;
}
__host__ __device__ __forceinline__ int rn_brick_fz(int z) { return (z >> 1) * 32 + (z & 1); }
__host__ __device__ __forceinline__ int rn_brick(const RnDev &p, int x, int y, int z) {
    return rn_brick_fx(p, x) + rn_brick_fy(p, y) + rn_brick_fz(z);
}

__host__ __device__ __forceinline__ void rn_unbrick(const RnDev &p, int off, int &x, int &y, int &z) {
    const int line = off >> 5, in = off & 31;
    const int lz = line % p.blz, t = line / p.blz;
    const int by = t % p.bby, bx = t / p.bby;
    x = bx * 4 + ((in >> 4) & 1) * 2 + ((in >> 2) & 1);
    y = by * 4 + ((in >> 3) & 1) * 2 + ((in >> 1) & 1);
    z = lz * 2 + (in & 1);
}

// ---- PTX: mbarrier + TMA bulk copies (1-D, no tensor map needed) ---------------------------
__device__ __forceinline__ uint32_t rn_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Canonical use: one thread initialises the CTA's barriers at kernel start, fences, and the
// whole CTA synchronises before anyone arms or waits on them.
// Cross-proxy hazard found by the parity tests on B200: a generic-proxy store (st.shared) issued
// right after try_wait succeeds, to bytes that the bulk copy has written, can lose against the
// async-proxy write.  The kernels therefore never store into a TMA-written range before they
// are done reading it (the zero tail of s_hat comes from global memory instead).
__device__ __forceinline__ void rn_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void rn_mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void rn_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rn_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void rn_bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void rn_bulk_s2g(void *dst, uint32_t src_smem, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(dst), "r"(src_smem), "r"(bytes), "l"(pol) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void rn_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void rn_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- fast math (tolerance-gated values only; integer decisions never come through here) ----
__device__ __forceinline__ float rn_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rn_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rn_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Occupancy-to-ray message (mrf_np.py:52-71) in a cancellation-free form.  With e = exp(-|x|),
// u = e / (1 + e) is min(o, 1 - o) to full relative precision; the clip of o to
// [1e-4, 1 - 1e-4] is max(u, 1e-4).  Returned as ONE signed float w: x >= 0 -> w = -u (o = 1 + w,
// 1 - o = -w), x < 0 -> w = u (o = w, 1 - o = 1 - w).  w == 0 marks "no voxel" (o = 0, 1 - o = 1).
__device__ __forceinline__ float rn_occ_w(float acc, float msg) {
    const float x = acc - msg;
    const float e = rn_ex2(-fabsf(x) * 1.4426950408889634f);
    const float u = fmaxf(e * rn_rcp(1.0f + e), 1e-4f);
    return (x >= 0.f) ? -u : u;
}
__device__ __forceinline__ void rn_occ_from_w(float w, float &o, float &q) {
    const bool pos = w >= 0.f;
    o = pos ? w : 1.0f + w;
    q = pos ? 1.0f - w : -w;
}

// =======================================================================================
// a1 + a3: DDA emitting bit-plane step codes, thread per ray
// =======================================================================================
struct DdaCodesArgs {
    const int32_t *ray_idxs;
    const float *P_inv, *centre;
    float *starts, *ends;      // optional outputs
    uint32_t *hdr;             // [n][2]
    uint8_t *codes;            // [n][code_stride]
    int32_t *count;            // [n]
    int64_t n_rays;
};

// ray_tracing.pyx:99-199 with the same operation order as rn_dda<> (rn_kernels.cuh); only the
// output differs.  Returns the voxel count.
__device__ __forceinline__ int rn_dda_codes(const RnDev &p, const float *rs_in, const float *re_in, uint32_t *hdr,
                                            uint2 *words) {
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
    hdr[0] = 0;
    hdr[1] = 0;
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
    hdr[0] = (uint32_t)cur[0] | ((uint32_t)cur[1] << 16);
    hdr[1] = (uint32_t)cur[2] | ((step[0] < 0 ? 1u : 0u) << 16) | ((step[1] < 0 ? 1u : 0u) << 17) |
             ((step[2] < 0 ? 1u : 0u) << 18);
    // The loop of ray_tracing.pyx:169-197 on counters instead of coordinates: togo_a = steps still to
    // take along a to stand on the last voxel (all three zero <=> cur == last; an axis that has
    // overshot goes negative and never returns to zero, like cur != last), rem_a = steps still
    // possible along a before leaving the grid (negative after the decrement <=> the stepped
    // coordinate left the grid: stop WITHOUT emitting).  The 32 voxels of one code word are unrolled,
    // so the bit position is an immediate and the word is stored once.
    int togo0 = (last[0] - cur[0]) * step[0], togo1 = (last[1] - cur[1]) * step[1], togo2 = (last[2] - cur[2]) * step[2];
    int rem0 = step[0] > 0 ? g[0] - 1 - cur[0] : cur[0], rem1 = step[1] > 0 ? g[1] - 1 - cur[1] : cur[1],
        rem2 = step[2] > 0 ? g[2] - 1 - cur[2] : cur[2];
    float tm0 = tMax[0], tm1 = tMax[1], tm2 = tMax[2];
    const float td0 = tDelta[0], td1 = tDelta[1], td2 = tDelta[2];
    const int M = p.M;
    int ii = 1;
    uint32_t lo = 1u, hi = 1u;   // voxel 0: "no step"
    bool finished = false;
    for (int w = 0; !finished; w++) {
#pragma unroll
        for (int pos = 0; pos < 32; pos++) {
            if (pos == 0 && w == 0) continue;
            if ((togo0 | togo1 | togo2) == 0 || ii >= M) { finished = true; break; }
            // strict '<': ties X=Y go to the Y/Z branch, any tie with Z goes to Z
            const bool xy = tm0 < tm1;
            const float tm = xy ? tm0 : tm1;
            const bool az = !(tm < tm2);
            const bool ax = xy && !az, ay = !xy && !az;
            togo0 -= ax ? 1 : 0; togo1 -= ay ? 1 : 0; togo2 -= az ? 1 : 0;
            rem0 -= ax ? 1 : 0; rem1 -= ay ? 1 : 0; rem2 -= az ? 1 : 0;
            if ((rem0 | rem1 | rem2) < 0) { finished = true; break; }
            tm0 = ax ? tm0 + td0 : tm0;
            tm1 = ay ? tm1 + td1 : tm1;
            tm2 = az ? tm2 + td2 : tm2;
            lo |= ay ? (1u << pos) : 0u;      // code = axis stepped along: x 0, y 1, z 2
            hi |= az ? (1u << pos) : 0u;
            ii++;
        }
        if (!finished) {
            words[w] = make_uint2(lo, hi);
            lo = 0u; hi = 0u;
        }
    }
    int nw = (ii + 31) >> 5;
    if (ii & 31) {   // flush the partial pair, tail padded with "no step"
        const uint32_t pad = ~((1u << (ii & 31)) - 1u);
        words[nw - 1] = make_uint2(lo | pad, hi | pad);
    }
    for (; (nw & 3) != 0; nw++) words[nw] = make_uint2(0xffffffffu, 0xffffffffu);   // pad to a whole chunk
    return ii;
}

__global__ void __launch_bounds__(128) dda_codes_kernel(RnDev p, DdaCodesArgs a) {
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
    a.count[r] = rn_dda_codes(p, rs, re, a.hdr + 2 * r, reinterpret_cast<uint2 *>(a.codes + r * (int64_t)p.code_stride));
}

// =======================================================================================
// Decoding: coordinates of voxel (32*sub + lane) from one (lo, hi) pair and the number of
// steps taken before the pair
// =======================================================================================
struct RayHead {
    int x0, y0, z0, sx, sy, sz;
};
__device__ __forceinline__ RayHead rn_ray_head(const uint32_t *hdr) {
    const uint32_t h0 = __ldg(hdr), h1 = __ldg(hdr + 1);
    RayHead d;
    d.x0 = h0 & 0xffff;
    d.y0 = h0 >> 16;
    d.z0 = h1 & 0xffff;
    d.sx = (h1 & (1u << 16)) ? -1 : 1;
    d.sy = (h1 & (1u << 17)) ? -1 : 1;
    d.sz = (h1 & (1u << 18)) ? -1 : 1;
    return d;
}
struct StepCount {
    int nx, ny, nz;
};
// Decode this lane's voxel of the pair; `before` is advanced past the pair (warp-uniform).
__device__ __forceinline__ void rn_decode_pair(const RayHead &h, uint32_t lo, uint32_t hi, int lane, StepCount &before,
                                               int &x, int &y, int &z) {
    const uint32_t mx = ~(hi | lo), my = ~hi & lo, mz = hi & ~lo;
    const uint32_t le = (2u << lane) - 1u;   // bits <= lane (lane 31: 0 - 1 = all ones)
    x = h.x0 + h.sx * (before.nx + __popc(mx & le));
    y = h.y0 + h.sy * (before.ny + __popc(my & le));
    z = h.z0 + h.sz * (before.nz + __popc(mz & le));
    before.nx += __popc(mx);
    before.ny += __popc(my);
    before.nz += __popc(mz);
}

// Coordinates of voxel i of a ray by walking its code words (one thread; used for the arg-max voxel).
__device__ __forceinline__ void rn_decode_single(const RayHead &h, const uint2 *words, int i, int &x, int &y, int &z) {
    int nx = 0, ny = 0, nz = 0;
    const int wl = i >> 5;
    for (int w = 0; w <= wl; w++) {
        const uint2 c = __ldg(words + w);
        const uint32_t m = (w < wl) ? 0xffffffffu : ((2u << (i & 31)) - 1u);
        nx += __popc(~(c.y | c.x) & m);
        ny += __popc(~c.y & c.x & m);
        nz += __popc(c.y & ~c.x & m);
    }
    x = h.x0 + h.sx * nx;
    y = h.y0 + h.sy * ny;
    z = h.z0 + h.sz * nz;
}

// =======================================================================================
// Ray binning: order[] groups rays by length class; inside a class rays follow a 2-D tiled
// enumeration of the image (8 x 8 pixel tiles) so that neighbouring warps work on
// neighbouring rays (shared voxels -> L1/L2 hits on the accumulator gathers)
// =======================================================================================
__device__ __forceinline__ int rn_class_of(int L) { return (L <= 1) ? 0 : ((L + RN_CHUNK - 1) / RN_CHUNK); }

// position t of the tiled enumeration -> ray position k (identity if seg_len <= 0).
// Pixels are enumerated in 8x8 tiles; when the image is made of whole 64x64 super-tiles the tiles
// are enumerated super-tile by super-tile, so that the few thousand rays in flight on the GPU
// at any time form a compact 2-D patch of the image: their epipolar bands in the other views
// (similarity kernel) and their voxels (BP gathers) then stay L2-resident.  A column-major walk
// of tiles would make the in-flight set a thin full-height strip whose epipolar fans cover most
// of every feature map.
__device__ __forceinline__ int64_t rn_tiled_position(int64_t t, int64_t seg_len, int H, int mode = 2) {
    if (seg_len <= 0 || mode == 0) return t;
    const int64_t seg = t / seg_len;
    const int tl = (int)(t - seg * seg_len);
    const int W = (int)(seg_len / H);
    int x, y;
    if (mode == 2 && (H & 63) == 0 && (W & 63) == 0) {
        const int st = tl >> 12, rem = tl & 4095;          // 64 tiles of 64 pixels
        const int sty = st % (H >> 6), stx = st / (H >> 6);
        const int tile = rem >> 6, in = rem & 63;
        x = stx * 64 + (tile >> 3) * 8 + (in >> 3);
        y = sty * 64 + (tile & 7) * 8 + (in & 7);
    } else if (mode == 3) {                                // tiles row-major: 8-pixel-tall strips
        const int tiles_x = W >> 3;
        const int tile = tl >> 6, in = tl & 63;
        x = (tile % tiles_x) * 8 + (in >> 3);
        y = (tile / tiles_x) * 8 + (in & 7);
    } else {
        const int tiles_y = H >> 3;
        const int tile = tl >> 6, in = tl & 63;
        x = (tile / tiles_y) * 8 + (in >> 3);
        y = (tile % tiles_y) * 8 + (in & 7);
    }
    return seg * seg_len + (int64_t)x * H + y;
}

__global__ void __launch_bounds__(256) bin_hist_kernel(const int32_t *count, int64_t n, unsigned long long *class_counts) {
    __shared__ unsigned int h[RN_NCLASS];
    if (threadIdx.x < RN_NCLASS) h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
        atomicAdd(&h[min(rn_class_of(__ldg(count + k)), RN_NCLASS - 1)], 1u);
    __syncthreads();
    if (threadIdx.x < RN_NCLASS && h[threadIdx.x]) atomicAdd(class_counts + threadIdx.x, (unsigned long long)h[threadIdx.x]);
}

// class_counts: [RN_NCLASS] totals (from bin_hist_kernel); cursors: [RN_NCLASS] zero-initialised.
__global__ void __launch_bounds__(256) bin_scatter_kernel(const int32_t *count, int64_t n, int64_t seg_len, int H,
                                                          const unsigned long long *class_counts,
                                                          unsigned long long *cursors, int32_t *order) {
    __shared__ unsigned int wcnt[8][RN_NCLASS];
    __shared__ unsigned long long base[RN_NCLASS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t k = -1;
    int cls = -1;
    if (t < n) {
        k = rn_tiled_position(t, seg_len, H);
        cls = min(rn_class_of(__ldg(count + k)), RN_NCLASS - 1);
    }
    unsigned int rank = 0;
#pragma unroll
    for (int c = 0; c < RN_NCLASS; c++) {
        const unsigned int m = __ballot_sync(RN_FULL_MASK, cls == c);
        if (cls == c) rank = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) wcnt[wid][c] = __popc(m);
    }
    __syncthreads();
    if (threadIdx.x < RN_NCLASS) {
        const int c = threadIdx.x;
        unsigned int tot = 0;
        for (int w = 0; w < 8; w++) { unsigned int v = wcnt[w][c]; wcnt[w][c] = tot; tot += v; }
        unsigned long long off = 0;
        for (int j = 0; j < c; j++) off += class_counts[j];
        base[c] = off + (tot ? atomicAdd(cursors + c, (unsigned long long)tot) : 0ull);
    }
    __syncthreads();
    if (cls >= 0) order[base[cls] + wcnt[wid][cls] + rank] = (int32_t)k;
}

// =======================================================================================
// a5 + a6: one BP sweep, warp per ray, state staged in shared memory
// =======================================================================================
struct Bp2Args {
    const int32_t *lin;     // resident layout: int32 [n][row_stride] bricked accumulator offset of every voxel
    const int32_t *idx;     // reference layout (kAos): int32 [n][M][3]
    const int32_t *count;
    const float *s_hat;     // resident: clip_and_renorm'ed rows; kAos: raw S_voxel_space rows
    float *msgs;
    const float *acc_in;    // resident: bricked; kAos: row-major [Gx][Gy][Gz]
    float *acc_out;
    const int32_t *order;   // null: identity
    int64_t first;          // first position of this launch inside order[]
    int64_t n;              // rays in this launch
    int nch_max;            // chunks of 128 voxels the shared-memory slots are sized for
    int rays_per_warp;      // bp4_kernel: consecutive rays one warp works through
    int uniform_acc;        // first sweep only: acc_in holds one value everywhere (the prior): no gathers needed
};

// bytes of dynamic shared memory one warp needs for rays of up to nch chunks:
//   sS [nch*128] f32  s_hat row, overwritten in place by cp_i * s_i
//   sM [nch*128] f32  message row, overwritten by w_i, then by the new messages
//   sLin[nch*128] i32 accumulator element offsets of the voxels
//   (8 words per chunk unused)
//   sPb [nch*32] f32  per-lane prefix base of every chunk;  sTot[pad4(nch)] chunk totals
//   sX  [128]    f32  transposition scratch
__host__ __device__ inline size_t rn_bp2_warp_bytes(int nch) {
    return sizeof(float) * ((size_t)nch * (128 * 3 + 8 + 32) + (size_t)((nch + 3) & ~3) + 128);
}
// kAos = the reference's buffers (voxel triplets, raw S clipped + renormalised on the fly as
// mrf_np.py:306 does, row-major accumulators, rows of any length M): same arithmetic, rows staged
// with ordinary loads because nothing guarantees the 16-byte alignment TMA needs.
//
// Voxels beyond the end of the ray inside its last 128-voxel chunk need no masks in the
// arithmetic: their s is 0 (so a_i = 0 and nothing reaches the sums), their message input is 0
// and their step code is "no step", so they alias the last real voxel; only the RED is predicated.
template <bool kFirst, bool kAos>
__global__ void __launch_bounds__(128) bp2_kernel(RnDev p, Bp2Args a) {
    extern __shared__ __align__(128) unsigned char rn_bp2_smem[];
    __shared__ __align__(8) uint64_t bars[4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nm = a.nch_max;

    if (!kAos) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int w = 0; w < 4; w++) rn_mbar_init(rn_smem_u32(&bars[w]), 1);
            rn_mbar_init_fence();
        }
        __syncthreads();
    }

    const int64_t k = (int64_t)blockIdx.x * 4 + wid;
    if (k >= a.n) return;
    const int64_t r = a.order ? (int64_t)__ldg(a.order + a.first + k) : a.first + k;
    const int L = __ldg(a.count + r);
    if (L <= 1) return;   // mrf_np.py:299-301
    const int nch = (L + RN_CHUNK - 1) / RN_CHUNK;

    float *sS = reinterpret_cast<float *>(rn_bp2_smem + (size_t)wid * rn_bp2_warp_bytes(nm));
    float *sM = sS + nm * 128;
    int *sLin = reinterpret_cast<int *>(sM + nm * 128);
    float *sPb = reinterpret_cast<float *>(sLin + nm * 128 + nm * 8);
    float *sTot = sPb + nm * 32;
    float *sX = sTot + ((nm + 3) & ~3);

    // ---- stage the ray's rows -----------------------------------------------------------------
    const uint32_t bar = rn_smem_u32(&bars[wid]);
    const int L4 = (L + 3) & ~3;
    const uint32_t row_bytes = (uint32_t)L4 * 4u;
    float *m_row = a.msgs + r * (int64_t)p.row_stride;
    const uint64_t pol_stream = rn_policy_evict_first();
    const uint64_t pol_keep = rn_policy_evict_last();
    float inv_raw = 1.f;
    const int32_t *idx_row = nullptr;
    if (kAos) {
        idx_row = a.idx + r * (int64_t)p.M * 3;
        const float *s_row = a.s_hat + r * (int64_t)p.row_stride;
        float part = 0.f;
        for (int i = lane; i < nch * RN_CHUNK; i += 32) {
            const bool ok = i < L;
            const float v = ok ? rn_clampf(s_row[i], 1e-5f, 0.99999f) : 0.f;   // mrf_np.py:4-8
            sS[i] = v;
            part += v;
            sM[i] = (!kFirst && ok) ? m_row[i] : 0.f;
        }
        inv_raw = 1.0f / rn_warp_sum(part);
        __syncwarp();
    } else {
        if (lane == 0) {
            rn_mbar_expect_tx(bar, row_bytes * (kFirst ? 2u : 3u));
            rn_bulk_g2s(rn_smem_u32(sLin), a.lin + r * (int64_t)p.row_stride, row_bytes, bar, pol_stream);
            rn_bulk_g2s(rn_smem_u32(sS), a.s_hat + r * (int64_t)p.row_stride, row_bytes, bar, pol_stream);
            if (!kFirst) rn_bulk_g2s(rn_smem_u32(sM), m_row, row_bytes, bar, pol_stream);
        }
        // s = 0 for the slots of the last chunk that the bulk copy does not cover (the front end
        // stores zeros in s_hat[L .. L4)).  Only s needs it: whatever message or accumulator value
        // a slot beyond the ray picks up, its a_i = o_i cp_i s_i is 0.  These generic-proxy stores
        // never touch a byte the async proxy writes.
        for (int i = L4 + lane; i < nch * RN_CHUNK; i += 32) sS[i] = 0.f;
        __syncwarp();
        rn_mbar_wait(bar, 0);
    }

    // ---- forward: gather, occupancy-to-ray, prefix scans ------------------------------------
    float ga[4];
    // issue the accumulator gathers of chunk c: lane-consecutive voxels share sectors.  Slots
    // beyond the ray (last chunk only) gather nothing; their value is irrelevant (s = 0).
    auto issue = [&](int c) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = c * RN_CHUNK + 32 * j + lane;
            if (kAos) {
                const int ii = min(i, L - 1);
                const int lin = rn_lin(p, __ldg(idx_row + 3 * ii), __ldg(idx_row + 3 * ii + 1), __ldg(idx_row + 3 * ii + 2));
                sLin[i] = lin;
                ga[j] = rn_ld_acc_pol(a.acc_in + lin, pol_keep);
            } else if (c < nch - 1 || i < L) {
                ga[j] = rn_ld_acc_pol(a.acc_in + sLin[i], pol_keep);
            } else {
                ga[j] = 0.f;
            }
        }
    };
    issue(0);
    float carry_cp = 1.f, carry_pre = 0.f;
    for (int c = 0; c < nch; c++) {
        // lane-consecutive -> 4 consecutive voxels per lane
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; j++) sX[32 * j + lane] = ga[j];
        __syncwarp();
        const float4 acc4 = *reinterpret_cast<const float4 *>(sX + 4 * lane);
        if (c + 1 < nch) issue(c + 1);   // next chunk's gathers fly while this one is computed
        const int i0 = c * RN_CHUNK + 4 * lane;
        const float4 s4 = *reinterpret_cast<const float4 *>(sS + i0);
        float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f);   // first sweep: messages are 0 (mrf_np.py:275)
        if (!kFirst || kAos) m4 = *reinterpret_cast<const float4 *>(sM + i0);
        const float accv[4] = {acc4.x, acc4.y, acc4.z, acc4.w};
        const float mv[4] = {m4.x, m4.y, m4.z, m4.w};
        float sv[4] = {s4.x, s4.y, s4.z, s4.w};
        float w[4], o[4], q[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            w[j] = rn_occ_w(accv[j], mv[j]);
            if (kAos) sv[j] *= inv_raw;
            rn_occ_from_w(w[j], o[j], q[j]);
        }
        // exclusive products cp_i = prod_{k<i} (1 - o_k)
        const float lp0 = q[0], lp1 = lp0 * q[1], lp2 = lp1 * q[2], lp3 = lp2 * q[3];
        const float inc = rn_warp_incl_scan_mul(lp3, lane);
        float exc = __shfl_up_sync(RN_FULL_MASK, inc, 1);
        if (lane == 0) exc = 1.f;
        const float base = carry_cp * exc;
        carry_cp = carry_cp * __shfl_sync(RN_FULL_MASK, inc, 31);
        float cps[4];
        cps[0] = base * sv[0];
        cps[1] = (base * lp0) * sv[1];
        cps[2] = (base * lp1) * sv[2];
        cps[3] = (base * lp2) * sv[3];
        // prefix sums of a_i = o_i cp_i s_i (true exclusive scan: no cancellation)
        const float la = fmaf(o[3], cps[3], fmaf(o[2], cps[2], fmaf(o[1], cps[1], o[0] * cps[0])));
        const float sinc = rn_warp_incl_scan_add(la, lane);
        float sexc = __shfl_up_sync(RN_FULL_MASK, sinc, 1);
        if (lane == 0) sexc = 0.f;
        const float tot = __shfl_sync(RN_FULL_MASK, sinc, 31);
        *reinterpret_cast<float4 *>(sM + i0) = make_float4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<float4 *>(sS + i0) = make_float4(cps[0], cps[1], cps[2], cps[3]);
        sPb[c * 32 + lane] = carry_pre + sexc;
        if (lane == 0) sTot[c] = tot;
        carry_pre += tot;
    }
    __syncwarp();

    // ---- backward: suffix sums, messages, scatter-add ----------------------------------------
    float carry_suf = 0.f;
    for (int c = nch - 1; c >= 0; c--) {
        const int i0 = c * RN_CHUNK + 4 * lane;
        const float4 w4 = *reinterpret_cast<const float4 *>(sM + i0);
        const float4 c4 = *reinterpret_cast<const float4 *>(sS + i0);
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
        const float cps[4] = {c4.x, c4.y, c4.z, c4.w};
        float o[4], q[4], av[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            rn_occ_from_w(w[j], o[j], q[j]);
            av[j] = o[j] * cps[j];
        }
        const float ra3 = av[3], ra2 = av[2] + ra3, ra1 = av[1] + ra2, ra0 = av[0] + ra1;
        // sum over the lanes ABOVE this one: shift, then inclusive reverse scan (exact exclusive)
        float above = __shfl_down_sync(RN_FULL_MASK, ra0, 1);
        if (lane == 31) above = 0.f;
        const float sbase = carry_suf + rn_warp_incl_rscan_add(above, lane);
        const float suf[4] = {sbase + ra1, sbase + ra2, sbase + ra3, sbase};
        float pre = sPb[c * 32 + lane];
        float msg[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float pos = pre + cps[j];
            const float neg = fmaf(suf[j], rn_rcp(q[j]), pre);
            // log p - log(1 - p) with p = pos / (pos + neg)
            msg[j] = 0.6931471805599453f * rn_lg2(pos * rn_rcp(neg));
            pre += av[j];
        }
        carry_suf += sTot[c];
        *reinterpret_cast<float4 *>(sM + i0) = make_float4(msg[0], msg[1], msg[2], msg[3]);
        __syncwarp();
        if (c < nch - 1) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = c * RN_CHUNK + 32 * j + lane;
                rn_red_add_pol(a.acc_out + sLin[i], sM[i], pol_keep);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = c * RN_CHUNK + 32 * j + lane;
                if (i < L) rn_red_add_pol(a.acc_out + sLin[i], sM[i], pol_keep);
            }
        }
    }
    if (kAos) {
        __syncwarp();
        for (int i = lane; i < L; i += 32) m_row[i] = sM[i];
        return;
    }
    // ---- new messages leave through one bulk store ----------------------------------------------
    rn_fence_async_smem();
    __syncwarp();
    if (lane == 0) {
        rn_bulk_s2g(m_row, rn_smem_u32(sM), row_bytes, pol_stream);
        rn_bulk_wait_read();
    }
    __syncwarp();
}

// =======================================================================================
// a8 + a9: depth re-estimation + arg-max -> depth, warp per ray, all images in one launch
// =======================================================================================
struct Depth2Args {
    const int32_t *lin;        // resident: int32 [n][row_stride] bricked offsets
    const int32_t *idx;        // kAos
    const int32_t *count;
    const float *s_hat;
    const float *msgs;
    const float *acc;          // bricked
    const float *axes;         // [Gx+Gy+Gz] voxel-centre coordinates per axis
    const float *centres;      // [n_seg][4] camera centres
    const int64_t *seg_starts; // [n_seg+1] first ray of every reference image (null: one image)
    int n_seg;
    float *depth_map;          // [n] optional
    float *S_new;              // optional [n][row_stride]: normalised depth distribution (tests)
    int64_t n_rays;
};

template <bool kAos>
__global__ void __launch_bounds__(128) depth2_kernel(RnDev p, Depth2Args a) {
    __shared__ __align__(16) float sXall[4][128];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * 4 + wid;
    if (r >= a.n_rays) return;
    float *sX = sXall[wid];
    const int L = __ldg(a.count + r);
    // reference image of this ray -> camera centre
    int lo_s = 0, hi_s = a.seg_starts ? a.n_seg : 1;
    while (hi_s - lo_s > 1) {
        const int mid = (lo_s + hi_s) >> 1;
        if (__ldg(a.seg_starts + mid) <= r) lo_s = mid; else hi_s = mid;
    }
    const float *C = a.centres ? a.centres + 4 * lo_s : nullptr;
    const int32_t *lin_row = nullptr;
    const int32_t *idx_row = nullptr;
    if (kAos) idx_row = a.idx + r * (int64_t)p.M * 3;
    else lin_row = a.lin + r * (int64_t)p.row_stride;
    const float *s_row = a.s_hat + r * (int64_t)p.row_stride;
    const float *m_row = a.msgs + r * (int64_t)p.row_stride;
    float rawsum = 1.f;
    if (kAos && L > 1) {   // clip_and_renorm on the fly (mrf_np.py:4-8, :379)
        float part = 0.f;
        for (int i = lane; i < L; i += 32) part += rn_clampf(s_row[i], 1e-5f, 0.99999f);
        rawsum = rn_warp_sum(part);
    }

    float bestv = -INFINITY;
    int besti = 0;
    float asum = 0.f;
    const uint64_t pol_keep = rn_policy_evict_last();
    if (L > 1) {   // mrf_np.py:376-377: rays with count <= 1 keep an all-zero row
        const int nch = (L + RN_CHUNK - 1) / RN_CHUNK;
        float carry_cp = 1.f;
        for (int c = 0; c < nch; c++) {
            float ga[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = c * RN_CHUNK + 32 * j + lane;
                int lin = 0;
                if (kAos) {
                    if (i < L) lin = rn_lin(p, __ldg(idx_row + 3 * i), __ldg(idx_row + 3 * i + 1), __ldg(idx_row + 3 * i + 2));
                } else if (i < L) {
                    lin = __ldg(lin_row + i);
                }
                ga[j] = (i < L) ? rn_ld_acc_pol(a.acc + lin, pol_keep) : 0.f;
            }
            const int i0 = c * RN_CHUNK + 4 * lane;
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), m4 = s4;
            if (kAos) {   // rows of arbitrary length M: no 16-byte alignment to rely on
                float ts[4], tm[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const bool ok = i0 + j < L;
                    ts[j] = ok ? rn_clampf(s_row[i0 + j], 1e-5f, 0.99999f) / rawsum : 0.f;
                    tm[j] = ok ? m_row[i0 + j] : 0.f;
                }
                s4 = make_float4(ts[0], ts[1], ts[2], ts[3]);
                m4 = make_float4(tm[0], tm[1], tm[2], tm[3]);
            } else if (i0 < L) {
                s4 = rn_ld_stream4(s_row + i0);
                m4 = rn_ld_stream4(m_row + i0);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; j++) sX[32 * j + lane] = ga[j];
            __syncwarp();
            const float4 acc4 = *reinterpret_cast<const float4 *>(sX + 4 * lane);
            const float accv[4] = {acc4.x, acc4.y, acc4.z, acc4.w};
            const float mv[4] = {m4.x, m4.y, m4.z, m4.w};
            float sv[4] = {s4.x, s4.y, s4.z, s4.w};
            float o[4], q[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool ok = i0 + j < L;
                const float w = ok ? rn_occ_w(accv[j], mv[j]) : 0.f;
                sv[j] = ok ? sv[j] : 0.f;
                rn_occ_from_w(w, o[j], q[j]);
            }
            const float lp0 = q[0], lp1 = lp0 * q[1], lp2 = lp1 * q[2], lp3 = lp2 * q[3];
            const float inc = rn_warp_incl_scan_mul(lp3, lane);
            float exc = __shfl_up_sync(RN_FULL_MASK, inc, 1);
            if (lane == 0) exc = 1.f;
            const float base = carry_cp * exc;
            carry_cp = carry_cp * __shfl_sync(RN_FULL_MASK, inc, 31);
            float av[4];
            av[0] = o[0] * (base * sv[0]);
            av[1] = o[1] * ((base * lp0) * sv[1]);
            av[2] = o[2] * ((base * lp1) * sv[2]);
            av[3] = o[3] * ((base * lp2) * sv[3]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (i0 + j < L) {
                    asum += av[j];
                    if (av[j] > bestv) { bestv = av[j]; besti = i0 + j; }
                }
            }
            if (a.S_new) {   // un-normalised for now; scaled below
                float *o_row = a.S_new + r * (int64_t)p.row_stride;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (i0 + j < L) o_row[i0 + j] = av[j];
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {   // first maximum over the ray (raynet_fp.py:193-205)
            const float ov = __shfl_xor_sync(RN_FULL_MASK, bestv, d);
            const int oi = __shfl_xor_sync(RN_FULL_MASK, besti, d);
            if (ov > bestv || (ov == bestv && oi < besti)) { bestv = ov; besti = oi; }
        }
        if (a.S_new) {
            asum = rn_warp_sum(asum);
            __syncwarp();
            float *o_row = a.S_new + r * (int64_t)p.row_stride;
            for (int i = lane; i < p.row_stride; i += 32) o_row[i] = (i < L) ? o_row[i] / asum : 0.f;
        }
    } else if (a.S_new) {
        float *o_row = a.S_new + r * (int64_t)p.row_stride;
        for (int i = lane; i < p.row_stride; i += 32) o_row[i] = 0.f;
    }
    if (lane == 0 && a.depth_map) {
        // voxel of the arg-max slot; an all-zero row selects slot 0, which holds the first voxel
        // of the ray or, for an empty ray, the zero-filled triplet (0, 0, 0) (raynet_fp.py:206-226)
        int x = 0, y = 0, z = 0;
        const int sel = (L > 1) ? besti : 0;
        if (kAos) {
            if (L >= 1) { x = idx_row[3 * sel]; y = idx_row[3 * sel + 1]; z = idx_row[3 * sel + 2]; }
        } else if (L >= 1) {
            rn_unbrick(p, __ldg(lin_row + sel), x, y, z);
        }
        const float cc[3] = {__ldg(a.axes + x), __ldg(a.axes + p.gx + y), __ldg(a.axes + p.gx + p.gy + z)};
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 3; i++) { const float dd = cc[i] - __ldg(C + i); sum += dd * dd; }
        a.depth_map[r] = sqrtf(sum);
    }
}

// =======================================================================================
// layout conversions + expansion of the resident state into the reference's buffers
// =======================================================================================
// row-major [Gx][Gy][Gz] -> bricks (padding voxels receive `pad`)
__global__ void grid_to_bricks_kernel(RnDev p, const float *grid, float *bricks, float pad, int64_t n_bricked) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_bricked; b += stride) {
        // invert the brick offset
        int x, y, z;
        rn_unbrick(p, (int)b, x, y, z);
        bricks[b] = (x < p.gx && y < p.gy && z < p.gz) ? grid[rn_lin(p, x, y, z)] : pad;
    }
}

// bricks -> row-major, optionally through the occupancy sigmoid (mrf_np.py:233-240)
__global__ void bricks_to_grid_kernel(RnDev p, const float *bricks, float *grid, int apply_sigmoid, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const int z = (int)(k % p.gz);
        const int64_t t = k / p.gz;
        const int y = (int)(t % p.gy), x = (int)(t / p.gy);
        float v = bricks[rn_brick(p, x, y, z)];
        if (apply_sigmoid) {
            const float e = expf(-fabsf(v));
            v = ((v >= 0.f) ? 1.0f : e) / (1.0f + e);
        }
        grid[k] = v;
    }
}

// Expand resident step codes into the reference's dense int32 [M][3] lists (thread per ray).
__global__ void expand_indices_kernel(RnDev p, const uint32_t *hdr, const uint8_t *codes, const int32_t *count,
                                      int32_t *idx, int64_t n_rays) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const int L = count[r];
    int32_t *row = idx + r * (int64_t)p.M * 3;
    const RayHead h = rn_ray_head(hdr + 2 * r);
    int x = h.x0, y = h.y0, z = h.z0;
    const uint2 *words = reinterpret_cast<const uint2 *>(codes + r * (int64_t)p.code_stride);
    uint2 cw = make_uint2(0, 0);
    for (int i = 0; i < p.M; i++) {
        if (i < L) {
            if ((i & 31) == 0) cw = words[i >> 5];
            const uint32_t f = (((cw.y >> (i & 31)) & 1u) << 1) | ((cw.x >> (i & 31)) & 1u);
            if (f == 0) x += h.sx;
            else if (f == 1) y += h.sy;
            else if (f == 2) z += h.sz;
            row[3 * i] = x; row[3 * i + 1] = y; row[3 * i + 2] = z;
        } else {
            row[3 * i] = 0; row[3 * i + 1] = 0; row[3 * i + 2] = 0;
        }
    }
}
