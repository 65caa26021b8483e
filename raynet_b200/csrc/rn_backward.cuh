// rn_backward.cuh -- SURVEY.md 8(f) row 3: the backward pass through the unrolled ray-potential BP.
//
// The reference trains through a TensorFlow graph (tf_implementations/forward_backward_pass.py:128-248):
//   scores -> softmax -> S (planes) -> single_ray_depth_to_voxels_map_li -> S_voxel_space
//          -> clip_and_renorm (mrf_tf.py:6-15) -> I unrolled BP sweeps (mrf_tf.py:60-143,176-249)
//          -> depth_estimate (mrf_tf.py:146-173,252-271) -> loss (loss_functions.py:4-35)
// and lets TF differentiate it.  Here the adjoint of every stage is a hand-written kernel over the
// reference's buffers (voxel lists int32 [N][M][3], rows float32 [N][M]); the forward sweeps are the
// ordinary kernels with the inputs of every sweep (messages, accumulator) kept as checkpoints.
//
// One ray, one sweep (s = S_norm, x_i = acc[v_i] - m_i, all sums over the ray's L voxels):
//   o_i = clip(sigmoid(x_i), 1e-4, 1-1e-4)   q_i = 1 - o_i   cp_i = prod_{k<i} q_k   c_i = cp_i s_i
//   a_i = o_i c_i   pre_i = sum_{j<i} a_j   suf_i = sum_{j>i} a_j
//   pos_i = c_i + pre_i   neg_i = pre_i + suf_i / q_i   m'_i = log pos_i - log neg_i
// Adjoint, g_i = dL/dm'_i (direct + gathered from the next accumulator's gradient):
//   gpos = g/pos   gneg = -g/neg   gpre = gpos + gneg   gsuf = gneg/q
//   ga_j = sum_{i>j} gpre_i + sum_{i<j} gsuf_i            (an exclusive suffix and an exclusive prefix sum)
//   gc_i = gpos_i + ga_i o_i      gs_i = gc_i cp_i        go_i = ga_i c_i - gq_i
//   gq_k = -gneg_k suf_k / q_k^2 + (sum_{i>k} gc_i s_i cp_i) / q_k
//   gx_i = go_i o_i (1 - o_i) where the sigmoid was not clipped, else 0
//   dL/dacc[v_i] += gx_i      dL/dm_i = -gx_i
//
// The sweep and depth adjoints run one WARP per ray (a lane per voxel, the recurrences as float64 warp scans with a
// carry between groups of 32 voxels), float64 arithmetic on the float32 checkpoints, per-voxel intermediates in a
// caller-owned float64 scratch laid out [ray][slot][voxel].  The front-end adjoint and the losses are still one
// thread per ray (a training batch is thousands of rays, not millions).
#pragma once

#include "rn_kernels.cuh"

#define RN_BWD_SLOTS 7      // float64 scratch values per voxel of rn_bp_sweep_backward

struct BwdArgs {
    const float *S;            // [n][M] raw S_voxel_space (clipped + renormalised here, mrf_tf.py:6-15)
    const int32_t *idx;        // [n][M][3]
    const int32_t *count;      // [n]
    const float *acc_in;       // [Gx][Gy][Gz] the accumulator the sweep (or the depth estimate) read
    const float *msg_in;       // [n][M] the messages it read (null: all zero, the first sweep)
    const float *g_out;        // sweep: [n][M] direct gradient w.r.t. its output messages (null: 0); depth: w.r.t. S_new
    const float *g_acc_next;   // sweep: [G] gradient w.r.t. the accumulator built from its output messages (null: 0)
    float *g_s;                // [n][M] += gradient w.r.t. S_norm
    float *g_msg_in;           // [n][M]  = gradient w.r.t. msg_in (may alias g_out)
    float *g_acc_in;           // [G]    += gradient w.r.t. acc_in (atomic)
    double *scratch;           // [n_chunk][slots][M]
    int64_t first, n;          // rays [first, first + n) of the arrays above
};

__device__ __forceinline__ bool rn_unclipped(double o) { return o > 1e-4 && o < 1 - 1e-4; }

// ---- float64 warp scans (a lane = a voxel of the current group of 32) ---------------------------------------------
__device__ __forceinline__ double rn_dscan_add(double v, int lane) {           // inclusive prefix sum
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_up_sync(RN_FULL_MASK, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ double rn_dscan_mul(double v, int lane) {           // inclusive prefix product
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_up_sync(RN_FULL_MASK, v, d);
        if (lane >= d) v *= t;
    }
    return v;
}
__device__ __forceinline__ double rn_drscan_add(double v, int lane) {          // inclusive suffix sum
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_down_sync(RN_FULL_MASK, v, d);
        if (lane + d < 32) v += t;
    }
    return v;
}
__device__ __forceinline__ double rn_dsum(double v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(RN_FULL_MASK, v, d);
    return v;
}
// exclusive versions from the inclusive scans: the neighbour's value, the identity at the open end
__device__ __forceinline__ double rn_dexcl_up(double inc, int lane, double identity) {
    const double t = __shfl_up_sync(RN_FULL_MASK, inc, 1);
    return lane == 0 ? identity : t;
}
__device__ __forceinline__ double rn_dexcl_down(double inc, int lane) {
    const double t = __shfl_down_sync(RN_FULL_MASK, inc, 1);
    return lane == 31 ? 0.0 : t;
}

// One warp per ray, a lane per voxel, groups of 32 voxels walked forwards (passes A, C) or backwards (B, D); the
// sequential recurrences of the derivation above become warp scans with a carry between groups.  Per-voxel
// intermediates live in the caller's float64 scratch, [ray][slot][voxel] (a warp reads and writes 256 contiguous bytes).
struct BwdWarp {
    const float *S_row;
    const int32_t *idx_row;
    const float *m_row;
    double Zs;
    int L;
    __device__ __forceinline__ void init(const RnDev &p, const BwdArgs &a, int64_t r, int lane) {
        L = a.count[r];
        S_row = a.S + r * (int64_t)p.M;
        idx_row = a.idx + r * (int64_t)p.M * 3;
        m_row = a.msg_in ? a.msg_in + r * (int64_t)p.M : nullptr;
        double z = 0.0;
        for (int i = lane; i < L; i += 32) z += (double)rn_clampf(S_row[i], 1e-5f, 0.99999f);
        Zs = rn_dsum(z);
    }
    __device__ __forceinline__ double s(int i) const { return (double)rn_clampf(S_row[i], 1e-5f, 0.99999f) / Zs; }
    __device__ __forceinline__ int vox(const RnDev &p, int i) const {
        return rn_lin(p, idx_row[3 * i], idx_row[3 * i + 1], idx_row[3 * i + 2]);
    }
    __device__ __forceinline__ double occ(const BwdArgs &a, int v, int i) const {
        const double x = (double)a.acc_in[v] - (m_row ? (double)m_row[i] : 0.0);
        const double e = exp(-fabs(x));
        const double o = ((x >= 0.0) ? 1.0 : e) / (1.0 + e);
        return fmin(fmax(o, 1e-4), 1 - 1e-4);
    }
};

__global__ void __launch_bounds__(128) bp_sweep_bwd_kernel(RnDev p, BwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t k = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (k >= a.n) return;
    const int64_t r = a.first + k;
    BwdWarp ray;
    ray.init(p, a, r, lane);
    const int L = ray.L;
    float *gm_row = a.g_msg_in + r * (int64_t)p.M;
    const float *go_row = a.g_out ? a.g_out + r * (int64_t)p.M : nullptr;
    if (L <= 1) {   // mrf_np.py:299-301 / the TF graph maps over count voxels: such rays produce no messages
        for (int i = lane; i < p.M; i += 32) gm_row[i] = 0.f;
        return;
    }
    double *scr = a.scratch + k * (int64_t)RN_BWD_SLOTS * p.M;
    auto sc = [&](int slot, int i) -> double & { return scr[slot * p.M + i]; };
    const int groups = (L + 31) >> 5;
    // pass A (forward): o, cp, pre
    double carry_cp = 1.0, carry_pre = 0.0;
    for (int gI = 0; gI < groups; gI++) {
        const int i = gI * 32 + lane;
        const bool ok = i < L;
        const double o = ok ? ray.occ(a, ray.vox(p, i), i) : 0.0;
        const double si = ok ? ray.s(i) : 0.0;
        const double qinc = rn_dscan_mul(1.0 - o, lane);
        const double cp = carry_cp * rn_dexcl_up(qinc, lane, 1.0);
        carry_cp *= __shfl_sync(RN_FULL_MASK, qinc, 31);
        const double ai = (o * cp) * si;
        const double ainc = rn_dscan_add(ai, lane);
        const double pre = carry_pre + rn_dexcl_up(ainc, lane, 0.0);
        carry_pre += __shfl_sync(RN_FULL_MASK, ainc, 31);
        if (ok) { sc(0, i) = o; sc(1, i) = cp; sc(2, i) = pre; }
    }
    // pass B (reverse): suffix sums, gpos / gneg, exclusive suffix sum of gpre
    double carry_suf = 0.0, carry_rs = 0.0;
    for (int gI = groups - 1; gI >= 0; gI--) {
        const int i = gI * 32 + lane;
        const bool ok = i < L;
        double o = 0.0, c = 0.0, pr = 0.0;
        if (ok) { o = sc(0, i); c = sc(1, i) * ray.s(i); pr = sc(2, i); }
        const double q = 1.0 - o;
        const double ainc = rn_drscan_add(o * c, lane);
        const double suf = carry_suf + rn_dexcl_down(ainc, lane);
        carry_suf += __shfl_sync(RN_FULL_MASK, ainc, 0);
        double gpos = 0.0, gneg = 0.0;
        if (ok) {
            const double pos = c + pr, neg = pr + suf / q;
            double g = go_row ? (double)go_row[i] : 0.0;
            if (a.g_acc_next) g += (double)a.g_acc_next[ray.vox(p, i)];
            gpos = g / pos; gneg = -g / neg;
        }
        const double ginc = rn_drscan_add(gpos + gneg, lane);
        const double rs = carry_rs + rn_dexcl_down(ginc, lane);          // sum_{j>i} gpre_j
        carry_rs += __shfl_sync(RN_FULL_MASK, ginc, 0);
        if (ok) {
            sc(3, i) = rs;
            sc(4, i) = gneg / q;                 // gsuf_i
            sc(5, i) = -gneg * suf / (q * q);    // direct part of gq_i
            sc(6, i) = gpos;
        }
    }
    // pass C (forward): ga, gs, go; h_i = gcp_i cp_i
    float *gs_row = a.g_s + r * (int64_t)p.M;
    double carry_ps = 0.0;
    for (int gI = 0; gI < groups; gI++) {
        const int i = gI * 32 + lane;
        const bool ok = i < L;
        const double gsuf = ok ? sc(4, i) : 0.0;
        const double pinc = rn_dscan_add(gsuf, lane);
        const double ps = carry_ps + rn_dexcl_up(pinc, lane, 0.0);       // sum_{j<i} gsuf_j
        carry_ps += __shfl_sync(RN_FULL_MASK, pinc, 31);
        if (ok) {
            const double o = sc(0, i), cpi = sc(1, i), si = ray.s(i), c = cpi * si;
            const double ga = sc(3, i) + ps;
            const double gc = sc(6, i) + ga * o;
            gs_row[i] += (float)(gc * cpi);
            sc(3, i) = (gc * si) * cpi;          // h_i
            sc(4, i) = ga * c;                   // go_i before the q term
        }
    }
    // pass D (reverse): gq, gx
    double carry_hs = 0.0;
    for (int gI = groups - 1; gI >= 0; gI--) {
        const int i = gI * 32 + lane;
        const bool ok = i < L;
        const double h = ok ? sc(3, i) : 0.0;
        const double hinc = rn_drscan_add(h, lane);
        const double hs = carry_hs + rn_dexcl_down(hinc, lane);          // sum_{j>i} h_j
        carry_hs += __shfl_sync(RN_FULL_MASK, hinc, 0);
        if (ok) {
            const double o = sc(0, i), q = 1 - o;
            const double gq = sc(5, i) + hs / q;
            const double go = sc(4, i) - gq;
            const double gx = rn_unclipped(o) ? go * o * (1 - o) : 0.0;
            gm_row[i] = (float)(-gx);
            if (gx != 0.0) atomicAdd(a.g_acc_in + ray.vox(p, i), (float)gx);
        }
    }
    for (int i = L + lane; i < p.M; i += 32) gm_row[i] = 0.f;
}

// depth_estimate (mrf_tf.py:146-173): P_i = a_i / sum_j a_j, gradient g_out w.r.t. P.  Warp per ray as above.
__global__ void __launch_bounds__(128) depth_bwd_kernel(RnDev p, BwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t k = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (k >= a.n) return;
    const int64_t r = a.first + k;
    BwdWarp ray;
    ray.init(p, a, r, lane);
    const int L = ray.L;
    float *gm_row = a.g_msg_in + r * (int64_t)p.M;
    const float *gP = a.g_out + r * (int64_t)p.M;
    if (L <= 1) {
        for (int i = lane; i < p.M; i += 32) gm_row[i] = 0.f;
        return;
    }
    double *scr = a.scratch + k * (int64_t)RN_BWD_SLOTS * p.M;
    auto sc = [&](int slot, int i) -> double & { return scr[slot * p.M + i]; };
    const int groups = (L + 31) >> 5;
    double carry_cp = 1.0, Zp = 0.0, dotp = 0.0;
    for (int gI = 0; gI < groups; gI++) {
        const int i = gI * 32 + lane;
        const bool ok = i < L;
        const double o = ok ? ray.occ(a, ray.vox(p, i), i) : 0.0;
        const double qinc = rn_dscan_mul(1.0 - o, lane);
        const double cp = carry_cp * rn_dexcl_up(qinc, lane, 1.0);
        carry_cp *= __shfl_sync(RN_FULL_MASK, qinc, 31);
        if (ok) {
            sc(0, i) = o; sc(1, i) = cp;
            const double ai = (o * cp) * ray.s(i);
            Zp += ai;
            dotp += (double)gP[i] * ai;
        }
    }
    const double Z = rn_dsum(Zp), dot = rn_dsum(dotp);
    float *gs_row = a.g_s + r * (int64_t)p.M;
    double carry_hs = 0.0;
    for (int gI = groups - 1; gI >= 0; gI--) {
        const int i = gI * 32 + lane;
        const bool ok = i < L;
        double o = 0.0, cpi = 0.0, si = 0.0, ga = 0.0, gc = 0.0;
        if (ok) {
            o = sc(0, i); cpi = sc(1, i); si = ray.s(i);
            ga = ((double)gP[i] - dot / Z) / Z;
            gc = ga * o;
            gs_row[i] += (float)(gc * cpi);
        }
        const double hinc = rn_drscan_add((gc * si) * cpi, lane);
        const double hs = carry_hs + rn_dexcl_down(hinc, lane);
        carry_hs += __shfl_sync(RN_FULL_MASK, hinc, 0);
        if (ok) {
            const double q = 1 - o;
            const double gq = hs / q;
            const double go = ga * (cpi * si) - gq;
            const double gx = rn_unclipped(o) ? go * o * (1 - o) : 0.0;
            gm_row[i] = (float)(-gx);
            if (gx != 0.0) atomicAdd(a.g_acc_in + ray.vox(p, i), (float)gx);
        }
    }
    for (int i = L + lane; i < p.M; i += 32) gm_row[i] = 0.f;
}

// Front-end adjoint: gradient w.r.t. S_norm -> clip_and_renorm (mrf_tf.py:6-15) -> normalised plane->voxel
// interpolation (planes_voxels_mapping.cu:6-92, same persistent two-pointer bracket as planes_to_voxels_kernel)
// -> plane distribution S [n][D] -> (optionally) softmax scores.  Thread per ray; the ray's D plane gradients
// are accumulated in its own row of g_S (zeroed here).
struct FrontBwdArgs {
    const float *axes;         // [Gx+Gy+Gz]
    const int32_t *idx, *count;
    const float *starts, *ends;   // [n][3]
    const float *S_planes;     // [n][D] the softmax output the forward pass interpolated
    const float *g_s_norm;     // [n][M] gradient w.r.t. S_norm (or w.r.t. S_voxel_space when g_is_raw)
    int g_is_raw;              // 1: the incoming gradient is already w.r.t. S_voxel_space (clip_and_renorm adjoint done)
    float *g_S_vox;            // [n][M] optional out: gradient w.r.t. S_voxel_space
    float *g_S;                // [n][D] out: gradient w.r.t. S_planes
    float *g_scores;           // [n][D] optional out: gradient w.r.t. the softmax input
    int64_t n;
};

// The bracket of one voxel: planes (right - 1, right) around t with the weights of planes_voxels_mapping.cu:54-92.
// The reference walks two persistent pointers along the ray; t never decreases along a ray (every DDA step moves the
// voxel centre along the ray's own direction), so the pointers stop at the smallest right >= 1 whose float test
// t - right * step <= 0 holds -- found here from ceil(t / step) with the SAME float expressions deciding.
struct Bracket {
    int left;
    double c1, c2;
};
__device__ __forceinline__ Bracket rn_bracket(const RnDev &p, const FrontBwdArgs &a, const int32_t *row, int i, const float *rs,
                                              const float *ray, float ray_norm, float step) {
    const float cc[3] = {a.axes[row[3 * i]], a.axes[p.gx + row[3 * i + 1]], a.axes[p.gx + p.gy + row[3 * i + 2]]};
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) { float vd = cc[j]; vd -= rs[j]; sum += ray[j] * vd; }
    const float t = rn_clampf(sum / ray_norm, 1e-4f, 1 - 1e-4f);
    int right = min(max((int)ceilf(t / step), 1), p.D - 1);
    while (right > 1 && t - (0.0f + (float)(right - 1) * step) <= 0) right--;
    while (right < p.D - 1 && t - (0.0f + (float)right * step) > 0) right++;
    const float left_d = fabsf(t - (0.0f + (float)(right - 1) * step)), right_d = fabsf(t - (0.0f + (float)right * step));
    Bracket b;
    b.left = right - 1;
    b.c1 = 1.0 - (double)(left_d / (left_d + right_d));
    b.c2 = 1.0 - (double)(right_d / (left_d + right_d));
    return b;
}

// Warp per ray, a lane per voxel, three passes over the ray (the sums Zu, then Zc and <g, S_norm>, then the gradients).
// The plane gradients of a ray are accumulated per warp in shared memory (float64): the lanes of a group whose voxels
// share a left plane form a run (left never decreases along the ray); the run is summed by a segmented warp scan and
// its last lane adds the two plane contributions.
__global__ void __launch_bounds__(128) frontend_bwd_kernel(RnDev p, FrontBwdArgs a) {
    __shared__ double sA[4][128], sW[4][128];     // per plane: sum_i w_ik gsv_i and sum_i w_ik (depth_planes <= 128)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * 4 + wid;
    if (r >= a.n) return;
    const int L = a.count[r];
    const int32_t *row = a.idx + r * (int64_t)p.M * 3;
    const float *Sr = a.S_planes + r * (int64_t)p.D;
    const float *gn = a.g_s_norm + r * (int64_t)p.M;
    float *gS = a.g_S + r * (int64_t)p.D;
    double *A = sA[wid], *Wt = sW[wid];
    for (int kpl = lane; kpl < p.D; kpl += 32) { A[kpl] = 0.0; Wt[kpl] = 0.0; }
    if (a.g_S_vox) for (int i = lane; i < p.M; i += 32) a.g_S_vox[r * (int64_t)p.M + i] = 0.f;
    __syncwarp();
    double Zu = 1.0, dot_v = 0.0;
    if (L > 0) {
        float rs[3], ray[3];
#pragma unroll
        for (int i = 0; i < 3; i++) { rs[i] = a.starts[3 * r + i]; ray[i] = a.ends[3 * r + i] - rs[i]; }
        float ray_norm = 0.f;
#pragma unroll
        for (int i = 0; i < 3; i++) ray_norm += ray[i] * ray[i];
        const float step = (1.0f - 0.0f) / (float)(p.D - 1);
        auto u_of = [&](int i, Bracket &b) -> double {
            b = rn_bracket(p, a, row, i, rs, ray, ray_norm, step);
            return b.c1 * (double)Sr[b.left] + b.c2 * (double)Sr[b.left + 1];
        };
        Bracket b;
        double z = 0.0;
        for (int i = lane; i < L; i += 32) z += u_of(i, b);
        Zu = rn_dsum(z);
        double zc = 0.0, dn = 0.0;
        for (int i = lane; i < L; i += 32) {
            const double sv = u_of(i, b) / Zu;                                  // S_voxel_space
            const double cl = fmin(fmax(sv, 1e-5), 1 - 1e-5);
            zc += cl; dn += (double)gn[i] * cl;
        }
        const double Zc = rn_dsum(zc), dot_n = rn_dsum(dn);
        const int groups = (L + 31) >> 5;
        for (int gI = 0; gI < groups; gI++) {
            const int i = gI * 32 + lane;
            const bool ok = i < L;
            double gsv = 0.0;
            b.left = p.D;   // lanes beyond the ray: a run of their own that writes nothing
            b.c1 = b.c2 = 0.0;
            if (ok) {
                const double sv = u_of(i, b) / Zu;
                // S_norm_i = cl_i / Zc;  dot_n / Zc = sum_j gn_j S_norm_j
                const double gcl = ((double)gn[i] - dot_n / Zc) / Zc;
                gsv = a.g_is_raw ? (double)gn[i] : ((sv > 1e-5 && sv < 1 - 1e-5) ? gcl : 0.0);
                if (a.g_S_vox) a.g_S_vox[r * (int64_t)p.M + i] = (float)gsv;
                dot_v += gsv * sv;
            }
            // segmented inclusive sums over runs of equal left: (c1 gsv, c2 gsv, c1, c2)
            double v0 = b.c1 * gsv, v1 = b.c2 * gsv, v2 = b.c1, v3 = b.c2;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int ol = __shfl_up_sync(RN_FULL_MASK, b.left, d);
                const double t0 = __shfl_up_sync(RN_FULL_MASK, v0, d), t1 = __shfl_up_sync(RN_FULL_MASK, v1, d);
                const double t2 = __shfl_up_sync(RN_FULL_MASK, v2, d), t3 = __shfl_up_sync(RN_FULL_MASK, v3, d);
                if (lane >= d && ol == b.left) { v0 += t0; v1 += t1; v2 += t2; v3 += t3; }
            }
            const int nl = __shfl_down_sync(RN_FULL_MASK, b.left, 1);
            const bool tail = ok && (lane == 31 || nl != b.left);
            if (tail) {   // (atomics: a right plane is the next run's left plane)
                atomicAdd(&A[b.left], v0); atomicAdd(&Wt[b.left], v2);
                atomicAdd(&A[b.left + 1], v1); atomicAdd(&Wt[b.left + 1], v3);
            }
        }
        dot_v = rn_dsum(dot_v);
    }
    __syncwarp();
    // gu_i = (gsv_i - sum_j gsv_j sv_j) / Zu, spread over the two planes of voxel i with its interpolation weights
    double dotS = 0.0;
    for (int kpl = lane; kpl < p.D; kpl += 32) {
        const double gk = (L > 0) ? (A[kpl] - dot_v * Wt[kpl]) / Zu : 0.0;
        A[kpl] = gk;
        gS[kpl] = (float)gk;
        dotS += gk * (double)Sr[kpl];
    }
    if (a.g_scores) {   // softmax: gscore_k = S_k (gS_k - sum_j gS_j S_j)
        dotS = rn_dsum(dotS);
        float *gz = a.g_scores + r * (int64_t)p.D;
        for (int kpl = lane; kpl < p.D; kpl += 32) gz[kpl] = (float)((double)Sr[kpl] * (A[kpl] - dotS));
    }
}

// gradient w.r.t. S_norm -> gradient w.r.t. the raw rows S (clip_and_renorm only), for callers that hand
// S_voxel_space to the MRF directly (mrf_tf.py:6-15; the reference's belief_propagation API).  Warp per ray.
__global__ void __launch_bounds__(128) clip_renorm_bwd_kernel(RnDev p, const float *S, const int32_t *count, const float *g_s_norm,
                                                              float *g_S, int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (r >= n) return;
    const int L = count[r];
    const float *Sr = S + r * (int64_t)p.M;
    const float *gn = g_s_norm + r * (int64_t)p.M;
    float *g = g_S + r * (int64_t)p.M;
    double zc = 0.0, dt = 0.0;
    for (int i = lane; i < L; i += 32) {
        const double cl = (double)rn_clampf(Sr[i], 1e-5f, 0.99999f);
        zc += cl;
        dt += (double)gn[i] * cl;
    }
    const double Zc = rn_dsum(zc), dot = rn_dsum(dt);
    for (int i = lane; i < p.M; i += 32) {
        float v = 0.f;
        if (i < L && L > 1) {
            const bool inside = Sr[i] > 1e-5f && Sr[i] < 0.99999f;
            if (inside) v = (float)(((double)gn[i] - dot / Zc) / Zc);
        }
        g[i] = v;
    }
}

// Losses of tf_implementations/loss_functions.py:4-35 on (S_target, S_pred) rows [n][M], value per ray and the
// gradient w.r.t. S_pred scaled by `scale` (1 / n for K.mean over the rays).
//   kind 0  emd:          mean_i |cumsum(y_true - y_pred)_i|   (mean over all M slots, like K.mean(axis=-1))
//   kind 1  squared_emd:  sum_i cumsum(y_true - y_pred)_i^2
//   kind 2  expected_squared_error: |sum_i (y_true_i - y_pred_i) dist_i|, dist_i = |centre(voxel_i) - camera centre|
struct LossArgs {
    const float *y_true, *y_pred;
    const int32_t *idx;        // kind 2
    const float *axes;         // kind 2
    const float *centres;      // kind 2: [n][4]
    float *loss;               // [n]
    float *g_pred;             // [n][M]
    float scale;
    int kind;
    int64_t n;
};

// Warp per ray: the cumulative sums as float64 warp scans over groups of 32 slots.
__global__ void __launch_bounds__(128) depth_loss_kernel(RnDev p, LossArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (r >= a.n) return;
    const float *yt = a.y_true + r * (int64_t)p.M, *yp = a.y_pred + r * (int64_t)p.M;
    float *g = a.g_pred + r * (int64_t)p.M;
    const int M = p.M;
    if (a.kind == 2) {
        const int32_t *row = a.idx + r * (int64_t)M * 3;
        const float *C = a.centres + 4 * r;
        double diff = 0.0;
        for (int i = lane; i < M; i += 32) {
            const float cc[3] = {a.axes[row[3 * i]], a.axes[p.gx + row[3 * i + 1]], a.axes[p.gx + p.gy + row[3 * i + 2]]};
            double d2 = 0.0;
            for (int j = 0; j < 3; j++) { const double dd = (double)cc[j] - (double)C[j]; d2 += dd * dd; }
            const double dist = sqrt(d2);
            diff += ((double)yt[i] - (double)yp[i]) * dist;
            g[i] = (float)dist;           // parked: the same lane rescales it below
        }
        diff = rn_dsum(diff);
        if (lane == 0) a.loss[r] = (float)fabs(diff);
        const double sg = (diff > 0.0) ? -1.0 : (diff < 0.0 ? 1.0 : 0.0);
        for (int i = lane; i < M; i += 32) g[i] = (float)(sg * (double)g[i] * (double)a.scale);
        return;
    }
    // c_i = cumsum(y_true - y_pred)_i;  dL/dy_pred_j = - sum_{i>=j} f'(c_i)
    const int groups = (M + 31) >> 5;
    double carry = 0.0, loss = 0.0;
    for (int gI = 0; gI < groups; gI++) {
        const int i = gI * 32 + lane;
        const bool ok = i < M;
        const double inc = rn_dscan_add(ok ? (double)yt[i] - (double)yp[i] : 0.0, lane);
        const double c = carry + inc;
        carry += __shfl_sync(RN_FULL_MASK, inc, 31);
        if (ok) {
            loss += (a.kind == 0) ? fabs(c) : c * c;
            g[i] = (float)c;           // parked for the reverse pass (read back by the same lane)
        }
    }
    loss = rn_dsum(loss);
    if (lane == 0) a.loss[r] = (float)((a.kind == 0) ? loss / M : loss);
    double tail = 0.0;
    for (int gI = groups - 1; gI >= 0; gI--) {
        const int i = gI * 32 + lane;
        const bool ok = i < M;
        double f = 0.0;
        if (ok) {
            const double ci = (double)g[i];
            f = (a.kind == 0) ? ((ci > 0.0) ? 1.0 : (ci < 0.0 ? -1.0 : 0.0)) / M : 2.0 * ci;
        }
        const double inc = rn_drscan_add(f, lane);           // sum over slots >= i of this group
        if (ok) g[i] = (float)(-(tail + inc) * (double)a.scale);
        tail += __shfl_sync(RN_FULL_MASK, inc, 0);
    }
}
