// rn_parity.cuh -- PARITY MODE of the BP sweep (a5 + a6) and of the depth pass (a8 + a9).
//
// BP amplifies a last-digit difference of the accumulator ~40x over five
// sweeps, so the reference's own float32- and float64-accumulator flavours disagree by 2e-5
// after a few sweeps and the end-to-end 1e-5 gate cannot be stated on float32 scatter-adds.
// mrf_np.py as it executes under NumPy >= 2 keeps both accumulators in float64
// (np.ones(f32) * np.float64, mrf_np.py:285-292), which makes the whole occupancy-to-ray chain
// float64 while messages, s and the pos / neg / p / log steps stay float32.  These kernels
// follow that flavour statement for statement:
//   * accumulators are float64 (gathered as doubles, scatter-added with RED.ADD.F64 -- the sum
//     over rays is then order-independent to ~1e-16);
//   * o_i, cp_i, a_i, prefix and suffix sums in float64 (mrf_np.py:52-112);
//   * pre32 = f32(prefix), pos = f32(pre32 + cp_i s_i), neg = f32(pre32 + suf_i / (1 - o_i)),
//     p = pos / (pos + neg) and logf(p) - logf(1 - p) in float32 (mrf_np.py:90-120).
// The suffix sum is total - prefix - a_i: in float64 the cancellation costs ~1e-16 absolute,
// five orders below anything the float32 steps after it can see.
//
// One warp per ray, one lane per voxel, 32 voxels per step; two passes over the ray (the first
// yields the total of the a_i, the second the messages) instead of per-voxel scratch.  This is
// the checking mode: ~4x the arithmetic of the fast kernels and 8-byte accumulator traffic.
#pragma once

#include "rn_engine.cuh"

struct ParityArgs {
    const int32_t *lin;        // resident layout: int32 [n][row_stride] bricked offsets
    const int32_t *idx;        // reference layout (kAos): int32 [n][M][3]
    const int32_t *count;
    const float *s;            // resident: s_hat rows; kAos: raw S_voxel_space rows (clipped + renormalised here)
    float *msgs;               // BP: in / out; depth: in
    const double *acc_in;      // resident: bricked; kAos: row-major [Gx][Gy][Gz]
    double *acc_out;           // BP only
    int first_sweep;           // messages count as 0 and are not read (mrf_np.py:275)
    float *S_new;              // depth: optional [n][row_stride]
    float *depth_map;          // depth: optional [n]
    const float *axes;         // depth: [Gx+Gy+Gz] voxel-centre coordinates
    const float *centres;      // depth: [n_seg][4]
    const int64_t *seg_starts; // depth: [n_seg + 1] or null
    int n_seg;
    int64_t n_rays;
};

__device__ __forceinline__ double rn_warp_sum_d(double v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(RN_FULL_MASK, v, d);
    return v;
}
__device__ __forceinline__ double rn_warp_incl_scan_add_d(double v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_up_sync(RN_FULL_MASK, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ double rn_warp_incl_scan_mul_d(double v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_up_sync(RN_FULL_MASK, v, d);
        if (lane >= d) v *= t;
    }
    return v;
}

// mrf_np.py:52-71 in float64: m = max(0, x); t1 = exp(0 - m); t2 = exp(x - m); o = t2 / (t2 + t1),
// clipped to [1e-4, 1 - 1e-4].  One of the two exponentials is exp(0) = 1 exactly.
__device__ __forceinline__ double rn_par_occ(double acc, float msg) {
    const double x = acc - (double)msg;
    const double e = exp(-fabs(x));
    const double t1 = (x >= 0.0) ? e : 1.0, t2 = (x >= 0.0) ? 1.0 : e;
    const double v = t2 / (t2 + t1);
    return fmin(fmax(v, 1e-4), 1 - 1e-4);
}

// Per-ray accessors for the two layouts.
template <bool kAos>
struct ParityRay {
    const int32_t *lin_row, *idx_row;
    const float *s_row;
    float *m_row;
    float fs;      // kAos: f32(sum of the clipped row) -- mrf_np.py:4-8
    int L;

    __device__ __forceinline__ void init(const RnDev &p, const ParityArgs &a, int64_t r, int lane) {
        L = __ldg(a.count + r);
        lin_row = kAos ? nullptr : a.lin + r * (int64_t)p.row_stride;
        idx_row = kAos ? a.idx + r * (int64_t)p.M * 3 : nullptr;
        s_row = a.s + r * (int64_t)p.row_stride;
        m_row = a.msgs + r * (int64_t)p.row_stride;
        fs = 1.f;
        if (kAos && L > 1) {
            double part = 0.0;
            for (int i = lane; i < L; i += 32) part += (double)rn_clampf(s_row[i], 1e-5f, 0.99999f);
            fs = (float)rn_warp_sum_d(part);
        }
    }
    __device__ __forceinline__ int offset(const RnDev &p, int i) const {
        if (kAos) return rn_lin(p, __ldg(idx_row + 3 * i), __ldg(idx_row + 3 * i + 1), __ldg(idx_row + 3 * i + 2));
        return __ldg(lin_row + i);
    }
    __device__ __forceinline__ float s(int i) const {
        return kAos ? rn_clampf(s_row[i], 1e-5f, 0.99999f) / fs : s_row[i];
    }
};

// One 32-voxel step of the forward chain: o_i, cp_i (exclusive product), a_i = (o_i cp_i) s_i, and the
// exclusive prefix sum of a.  Carries are warp-uniform.
template <bool kAos>
__device__ __forceinline__ void rn_par_step(const RnDev &p, const ParityArgs &a, const ParityRay<kAos> &ray, int i,
                                            int lane, bool read_msgs, double &carry_cp, double &carry_pre, int &off,
                                            float &si, double &o, double &cp, double &ai, double &pre) {
    const bool ok = i < ray.L;
    o = 0.0; si = 0.f; off = 0;
    double q = 1.0;
    if (ok) {
        off = ray.offset(p, i);
        const float m = read_msgs ? ray.m_row[i] : 0.f;
        o = rn_par_occ(a.acc_in[off], m);
        q = 1 - o;
        si = ray.s(i);
    }
    const double inc = rn_warp_incl_scan_mul_d(q, lane);
    double exc = __shfl_up_sync(RN_FULL_MASK, inc, 1);
    if (lane == 0) exc = 1.0;
    cp = carry_cp * exc;
    carry_cp = carry_cp * __shfl_sync(RN_FULL_MASK, inc, 31);
    ai = (o * cp) * (double)si;      // mrf_np.py:91
    const double sinc = rn_warp_incl_scan_add_d(ai, lane);
    double sexc = __shfl_up_sync(RN_FULL_MASK, sinc, 1);
    if (lane == 0) sexc = 0.0;
    pre = carry_pre + sexc;
    carry_pre = carry_pre + __shfl_sync(RN_FULL_MASK, sinc, 31);
}

template <bool kAos>
__global__ void __launch_bounds__(128) bp_parity_kernel(RnDev p, ParityArgs a) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * 4 + wid;
    if (r >= a.n_rays) return;
    ParityRay<kAos> ray;
    ray.init(p, a, r, lane);
    const int L = ray.L;
    if (L <= 1) return;   // mrf_np.py:299-301
    const bool read_msgs = a.first_sweep == 0;
    int off;
    float si;
    double o, cp, ai, pre;
    // pass 1: total of the a_i
    double carry_cp = 1.0, total = 0.0;
    for (int c0 = 0; c0 < L; c0 += 32)
        rn_par_step<kAos>(p, a, ray, c0 + lane, lane, read_msgs, carry_cp, total, off, si, o, cp, ai, pre);
    // pass 2: messages (identical arithmetic, so the prefixes are the ones total was built from)
    carry_cp = 1.0;
    double carry_pre = 0.0;
    for (int c0 = 0; c0 < L; c0 += 32) {
        const int i = c0 + lane;
        rn_par_step<kAos>(p, a, ray, i, lane, read_msgs, carry_cp, carry_pre, off, si, o, cp, ai, pre);
        if (i < L) {
            const double suf = fmax(0.0, (total - pre) - ai);
            const float pre32 = (float)pre;                                   // mrf_np.py:90-92
            const float pos = (float)((double)pre32 + cp * (double)si);       // :95
            const float neg = (float)((double)pre32 + suf / (1 - o));         // :109-112
            const float pr = pos / (pos + neg);                               // :115-116
            const float t = logf(pr) - logf(1 - pr);                          // :120
            ray.m_row[i] = t;
            atomicAdd(a.acc_out + off, (double)t);
        }
    }
}

// mrf_np.py:129-203 + :333-385 (+ raynet_fp.py:193-226 for the depth map) in the same flavour.
template <bool kAos>
__global__ void __launch_bounds__(128) depth_parity_kernel(RnDev p, ParityArgs a) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * 4 + wid;
    if (r >= a.n_rays) return;
    ParityRay<kAos> ray;
    ray.init(p, a, r, lane);
    const int L = ray.L;
    float *o_row = a.S_new ? a.S_new + r * (int64_t)p.row_stride : nullptr;
    float bestv = -INFINITY;
    int besti = 0;
    if (L > 1) {
        int off;
        float si;
        double o, cp, ai, pre;
        double carry_cp = 1.0, total = 0.0;
        for (int c0 = 0; c0 < L; c0 += 32)
            rn_par_step<kAos>(p, a, ray, c0 + lane, lane, true, carry_cp, total, off, si, o, cp, ai, pre);
        carry_cp = 1.0;
        double carry_pre = 0.0;
        for (int c0 = 0; c0 < p.row_stride; c0 += 32) {
            const int i = c0 + lane;
            float v = 0.f;
            if (c0 < L) {
                rn_par_step<kAos>(p, a, ray, i, lane, true, carry_cp, carry_pre, off, si, o, cp, ai, pre);
                if (i < L) {
                    v = (float)(ai / total);                                  // P / P.sum(), stored float32 (:378)
                    if (v > bestv) { bestv = v; besti = i; }
                }
            }
            if (o_row && i < p.row_stride) o_row[i] = v;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {   // first maximum over the ray (raynet_fp.py:193-205)
            const float ov = __shfl_xor_sync(RN_FULL_MASK, bestv, d);
            const int oi = __shfl_xor_sync(RN_FULL_MASK, besti, d);
            if (ov > bestv || (ov == bestv && oi < besti)) { bestv = ov; besti = oi; }
        }
    } else if (o_row) {
        for (int i = lane; i < p.row_stride; i += 32) o_row[i] = 0.f;
    }
    if (lane == 0 && a.depth_map) {
        int x = 0, y = 0, z = 0;
        const int sel = (L > 1) ? besti : 0;
        if (L >= 1) {
            if (kAos) { x = ray.idx_row[3 * sel]; y = ray.idx_row[3 * sel + 1]; z = ray.idx_row[3 * sel + 2]; }
            else rn_unbrick(p, __ldg(ray.lin_row + sel), x, y, z);
        }
        int seg = 0;
        if (a.seg_starts) {
            int lo_s = 0, hi_s = a.n_seg;
            while (hi_s - lo_s > 1) {
                const int mid = (lo_s + hi_s) >> 1;
                if (__ldg(a.seg_starts + mid) <= r) lo_s = mid; else hi_s = mid;
            }
            seg = lo_s;
        }
        const float *C = a.centres + 4 * seg;
        const float cc[3] = {__ldg(a.axes + x), __ldg(a.axes + p.gx + y), __ldg(a.axes + p.gx + p.gy + z)};
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 3; i++) { const float dd = cc[i] - __ldg(C + i); sum += dd * dd; }
        a.depth_map[r] = sqrtf(sum);
    }
}

// ---- float64 grids: fill, layout conversion, occupancy -------------------------------------------------
__global__ void fill_f64_kernel(double *dst, double value, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = value;
}

__global__ void grid_to_bricks_f64_kernel(RnDev p, const double *grid, double *bricks, double pad, int64_t n_bricked) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_bricked; b += stride) {
        int x, y, z;
        rn_unbrick(p, (int)b, x, y, z);
        bricks[b] = (x < p.gx && y < p.gy && z < p.gz) ? grid[rn_lin(p, x, y, z)] : pad;
    }
}

// mrf_np.py:233-240 in float64, stored float32
__device__ __forceinline__ float rn_sigmoid_f64(double x) {
    const double e = exp(-fabs(x));
    return (float)(((x >= 0.0) ? 1.0 : e) / (1.0 + e));
}

// bricks -> row-major float64 grid (optional) and / or float32 occupancy sigmoid(acc) (optional)
__global__ void bricks_to_grid_f64_kernel(RnDev p, const double *bricks, double *grid, float *occ, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const int z = (int)(k % p.gz);
        const int64_t t = k / p.gz;
        const int y = (int)(t % p.gy), x = (int)(t / p.gy);
        const double v = bricks[rn_brick(p, x, y, z)];
        if (grid) grid[k] = v;
        if (occ) occ[k] = rn_sigmoid_f64(v);
    }
}

__global__ void occupancy_f64_kernel(const double *acc, float *out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = rn_sigmoid_f64(acc[i]);
}
