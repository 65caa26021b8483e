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
// Thread per ray, sequential passes along the ray, float64 arithmetic on the float32 checkpoints, per-voxel
// intermediates in a caller-owned float64 scratch laid out [slot][voxel][ray] (coalesced across the rays of
// a launch).  A training batch is thousands of rays, not millions; this is the correct-first version.
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
    double *scratch;           // [slots][M][n_chunk]
    int64_t first, n;          // rays [first, first + n) of the arrays above
};

struct BwdRay {
    const float *S_row;
    const int32_t *idx_row;
    const float *m_row;
    double Zs;
    int L;
    __device__ __forceinline__ void init(const RnDev &p, const BwdArgs &a, int64_t r) {
        L = a.count[r];
        S_row = a.S + r * (int64_t)p.M;
        idx_row = a.idx + r * (int64_t)p.M * 3;
        m_row = a.msg_in ? a.msg_in + r * (int64_t)p.M : nullptr;
        Zs = 0.0;
        for (int i = 0; i < L; i++) Zs += (double)rn_clampf(S_row[i], 1e-5f, 0.99999f);
    }
    __device__ __forceinline__ double s(int i) const { return (double)rn_clampf(S_row[i], 1e-5f, 0.99999f) / Zs; }
    __device__ __forceinline__ int vox(const RnDev &p, int i) const {
        return rn_lin(p, idx_row[3 * i], idx_row[3 * i + 1], idx_row[3 * i + 2]);
    }
    __device__ __forceinline__ double occ(const RnDev &p, const BwdArgs &a, int i) const {
        const double x = (double)a.acc_in[vox(p, i)] - (m_row ? (double)m_row[i] : 0.0);
        const double e = exp(-fabs(x));
        const double v = ((x >= 0.0) ? 1.0 : e) / (1.0 + e);
        return fmin(fmax(v, 1e-4), 1 - 1e-4);
    }
};
__device__ __forceinline__ bool rn_unclipped(double o) { return o > 1e-4 && o < 1 - 1e-4; }

__global__ void __launch_bounds__(128) bp_sweep_bwd_kernel(RnDev p, BwdArgs a) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n) return;
    const int64_t r = a.first + k;
    BwdRay ray;
    ray.init(p, a, r);
    const int L = ray.L;
    float *gm_row = a.g_msg_in + r * (int64_t)p.M;
    const float *go_row = a.g_out ? a.g_out + r * (int64_t)p.M : nullptr;
    if (L <= 1) {   // mrf_np.py:299-301 / the TF graph maps over count voxels: such rays produce no messages
        for (int i = 0; i < p.M; i++) gm_row[i] = 0.f;
        return;
    }
    const int64_t plane = (int64_t)p.M * a.n;
    auto sc = [&](int slot, int i) -> double & { return a.scratch[slot * plane + (int64_t)i * a.n + k]; };
    // pass A (forward): o, cp, pre
    double cp = 1.0, pre = 0.0;
    for (int i = 0; i < L; i++) {
        const double o = ray.occ(p, a, i);
        sc(0, i) = o; sc(1, i) = cp; sc(2, i) = pre;
        pre += (o * cp) * ray.s(i);
        cp *= 1 - o;
    }
    // pass B (reverse): suffix sums, gpos / gneg, exclusive suffix sum of gpre
    double suf = 0.0, rs_gpre = 0.0;
    for (int i = L - 1; i >= 0; i--) {
        const double o = sc(0, i), q = 1 - o, c = sc(1, i) * ray.s(i), pr = sc(2, i);
        const double pos = c + pr, neg = pr + suf / q;
        double g = go_row ? (double)go_row[i] : 0.0;
        if (a.g_acc_next) g += (double)a.g_acc_next[ray.vox(p, i)];
        const double gpos = g / pos, gneg = -g / neg;
        sc(3, i) = rs_gpre;                  // sum_{j>i} gpre_j
        sc(4, i) = gneg / q;                 // gsuf_i
        sc(5, i) = -gneg * suf / (q * q);    // direct part of gq_i
        sc(6, i) = gpos;
        rs_gpre += gpos + gneg;
        suf += o * c;
    }
    // pass C (forward): ga, gs, go; h_i = gcp_i cp_i
    float *gs_row = a.g_s + r * (int64_t)p.M;
    double ps_gsuf = 0.0;
    for (int i = 0; i < L; i++) {
        const double o = sc(0, i), cpi = sc(1, i), si = ray.s(i), c = cpi * si;
        const double ga = sc(3, i) + ps_gsuf;
        ps_gsuf += sc(4, i);
        const double gc = sc(6, i) + ga * o;
        gs_row[i] += (float)(gc * cpi);
        sc(3, i) = (gc * si) * cpi;          // h_i
        sc(4, i) = ga * c;                   // go_i before the q term
    }
    // pass D (reverse): gq, gx
    double hs = 0.0;
    for (int i = L - 1; i >= 0; i--) {
        const double o = sc(0, i), q = 1 - o;
        const double gq = sc(5, i) + hs / q;
        hs += sc(3, i);
        const double go = sc(4, i) - gq;
        const double gx = rn_unclipped(o) ? go * o * (1 - o) : 0.0;
        gm_row[i] = (float)(-gx);
        if (gx != 0.0) atomicAdd(a.g_acc_in + ray.vox(p, i), (float)gx);
    }
    for (int i = L; i < p.M; i++) gm_row[i] = 0.f;
}

// depth_estimate (mrf_tf.py:146-173): P_i = a_i / sum_j a_j, gradient g_out w.r.t. P.
__global__ void __launch_bounds__(128) depth_bwd_kernel(RnDev p, BwdArgs a) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n) return;
    const int64_t r = a.first + k;
    BwdRay ray;
    ray.init(p, a, r);
    const int L = ray.L;
    float *gm_row = a.g_msg_in + r * (int64_t)p.M;
    const float *gP = a.g_out + r * (int64_t)p.M;
    if (L <= 1) {
        for (int i = 0; i < p.M; i++) gm_row[i] = 0.f;
        return;
    }
    const int64_t plane = (int64_t)p.M * a.n;
    auto sc = [&](int slot, int i) -> double & { return a.scratch[slot * plane + (int64_t)i * a.n + k]; };
    double cp = 1.0, Z = 0.0, dot = 0.0;
    for (int i = 0; i < L; i++) {
        const double o = ray.occ(p, a, i);
        sc(0, i) = o; sc(1, i) = cp;
        const double ai = (o * cp) * ray.s(i);
        Z += ai;
        dot += (double)gP[i] * ai;
        cp *= 1 - o;
    }
    float *gs_row = a.g_s + r * (int64_t)p.M;
    double hs = 0.0;
    for (int i = L - 1; i >= 0; i--) {
        const double o = sc(0, i), q = 1 - o, cpi = sc(1, i), si = ray.s(i);
        const double ga = ((double)gP[i] - dot / Z) / Z;
        const double gc = ga * o;
        gs_row[i] += (float)(gc * cpi);
        const double gq = hs / q;
        hs += (gc * si) * cpi;
        const double go = ga * (cpi * si) - gq;
        const double gx = rn_unclipped(o) ? go * o * (1 - o) : 0.0;
        gm_row[i] = (float)(-gx);
        if (gx != 0.0) atomicAdd(a.g_acc_in + ray.vox(p, i), (float)gx);
    }
    for (int i = L; i < p.M; i++) gm_row[i] = 0.f;
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

__global__ void __launch_bounds__(128) frontend_bwd_kernel(RnDev p, FrontBwdArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    const int L = a.count[r];
    const int32_t *row = a.idx + r * (int64_t)p.M * 3;
    const float *Sr = a.S_planes + r * (int64_t)p.D;
    const float *gn = a.g_s_norm + r * (int64_t)p.M;
    float *gS = a.g_S + r * (int64_t)p.D;
    for (int kpl = 0; kpl < p.D; kpl++) gS[kpl] = 0.f;
    if (a.g_S_vox) for (int i = 0; i < p.M; i++) a.g_S_vox[r * (int64_t)p.M + i] = 0.f;
    if (L > 0) {
        float rs[3], ray[3];
        for (int i = 0; i < 3; i++) { rs[i] = a.starts[3 * r + i]; ray[i] = a.ends[3 * r + i] - rs[i]; }
        float ray_norm = 0.f;
        for (int i = 0; i < 3; i++) ray_norm += ray[i] * ray[i];
        const float step = (1.0f - 0.0f) / (float)(p.D - 1);
        // the forward interpolation, twice: first for the sums, then for the gradients
        double Zu = 0.0, Zc = 0.0, dot_n = 0.0, dot_v = 0.0;
        for (int pass = 0; pass < 3; pass++) {
            int left = 0, right = 1;
            for (int i = 0; i < L; i++) {
                const float cc[3] = {a.axes[row[3 * i]], a.axes[p.gx + row[3 * i + 1]], a.axes[p.gx + p.gy + row[3 * i + 2]]};
                float sum = 0.f;
                for (int j = 0; j < 3; j++) { float vd = cc[j]; vd -= rs[j]; sum += ray[j] * vd; }
                const float t = rn_clampf(sum / ray_norm, 1e-4f, 1 - 1e-4f);
                float left_d = t - (0.0f + (float)left * step), right_d = t - (0.0f + (float)right * step);
                while (left_d > 0 && right_d > 0) {
                    left++; right++;
                    left_d = t - (0.0f + (float)left * step);
                    right_d = t - (0.0f + (float)right * step);
                }
                left_d = fabsf(left_d); right_d = fabsf(right_d);
                const double c1 = 1.0 - (double)(left_d / (left_d + right_d));
                const double c2 = 1.0 - (double)(right_d / (left_d + right_d));
                const double u = c1 * (double)Sr[left] + c2 * (double)Sr[right];
                if (pass == 0) { Zu += u; continue; }
                const double sv = u / Zu;                                  // S_voxel_space
                const double cl = fmin(fmax(sv, 1e-5), 1 - 1e-5);
                if (pass == 1) { Zc += cl; dot_n += (double)gn[i] * cl; continue; }
                // pass 2: S_norm_i = cl_i / Zc;  dot_n / Zc = sum_j gn_j S_norm_j
                const double gcl = ((double)gn[i] - dot_n / Zc) / Zc;
                const double gsv = a.g_is_raw ? (double)gn[i] : ((sv > 1e-5 && sv < 1 - 1e-5) ? gcl : 0.0);
                if (a.g_S_vox) a.g_S_vox[r * (int64_t)p.M + i] = (float)gsv;
                dot_v += gsv * sv;
                // gu_i = (gsv_i - sum_j gsv_j sv_j) / Zu needs the complete dot_v: accumulate the two parts
                gS[left] += (float)(c1 * gsv / Zu);
                gS[right] += (float)(c2 * gsv / Zu);
            }
        }
        // the - (sum_j gsv_j sv_j) / Zu part: sum_i c1_i dS[l_i] + c2_i dS[r_i] = du_i, and sum_i du_i weights are the
        // interpolation weights again; d(sum u)/dS_k = W_k with sum_k W_k S_k = Zu.  Apply it with a fourth sweep.
        int left = 0, right = 1;
        for (int i = 0; i < L; i++) {
            const float cc[3] = {a.axes[row[3 * i]], a.axes[p.gx + row[3 * i + 1]], a.axes[p.gx + p.gy + row[3 * i + 2]]};
            float sum = 0.f;
            for (int j = 0; j < 3; j++) { float vd = cc[j]; vd -= rs[j]; sum += ray[j] * vd; }
            const float t = rn_clampf(sum / ray_norm, 1e-4f, 1 - 1e-4f);
            float left_d = t - (0.0f + (float)left * step), right_d = t - (0.0f + (float)right * step);
            while (left_d > 0 && right_d > 0) {
                left++; right++;
                left_d = t - (0.0f + (float)left * step);
                right_d = t - (0.0f + (float)right * step);
            }
            left_d = fabsf(left_d); right_d = fabsf(right_d);
            const double c1 = 1.0 - (double)(left_d / (left_d + right_d));
            const double c2 = 1.0 - (double)(right_d / (left_d + right_d));
            gS[left] -= (float)(c1 * dot_v / Zu);
            gS[right] -= (float)(c2 * dot_v / Zu);
        }
    }
    if (a.g_scores) {   // softmax: gscore_k = S_k (gS_k - sum_j gS_j S_j)
        float *gz = a.g_scores + r * (int64_t)p.D;
        double dot = 0.0;
        for (int kpl = 0; kpl < p.D; kpl++) dot += (double)gS[kpl] * (double)Sr[kpl];
        for (int kpl = 0; kpl < p.D; kpl++) gz[kpl] = (float)((double)Sr[kpl] * ((double)gS[kpl] - dot));
    }
}

// gradient w.r.t. S_norm -> gradient w.r.t. the raw rows S (clip_and_renorm only), for callers that hand
// S_voxel_space to the MRF directly (mrf_tf.py:6-15; the reference's belief_propagation API)
__global__ void __launch_bounds__(128) clip_renorm_bwd_kernel(RnDev p, const float *S, const int32_t *count, const float *g_s_norm,
                                                              float *g_S, int64_t n) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int L = count[r];
    const float *Sr = S + r * (int64_t)p.M;
    const float *gn = g_s_norm + r * (int64_t)p.M;
    float *g = g_S + r * (int64_t)p.M;
    double Zc = 0.0, dot = 0.0;
    for (int i = 0; i < L; i++) {
        const double cl = (double)rn_clampf(Sr[i], 1e-5f, 0.99999f);
        Zc += cl;
        dot += (double)gn[i] * cl;
    }
    for (int i = 0; i < p.M; i++) {
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

__global__ void __launch_bounds__(128) depth_loss_kernel(RnDev p, LossArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    const float *yt = a.y_true + r * (int64_t)p.M, *yp = a.y_pred + r * (int64_t)p.M;
    float *g = a.g_pred + r * (int64_t)p.M;
    const int M = p.M;
    if (a.kind == 2) {
        const int32_t *row = a.idx + r * (int64_t)M * 3;
        const float *C = a.centres + 4 * r;
        double diff = 0.0;
        for (int i = 0; i < M; i++) {
            const float cc[3] = {a.axes[row[3 * i]], a.axes[p.gx + row[3 * i + 1]], a.axes[p.gx + p.gy + row[3 * i + 2]]};
            double d2 = 0.0;
            for (int j = 0; j < 3; j++) { const double dd = (double)cc[j] - (double)C[j]; d2 += dd * dd; }
            const double dist = sqrt(d2);
            diff += ((double)yt[i] - (double)yp[i]) * dist;
            g[i] = (float)dist;
        }
        a.loss[r] = (float)fabs(diff);
        const double sg = (diff > 0.0) ? -1.0 : (diff < 0.0 ? 1.0 : 0.0);
        for (int i = 0; i < M; i++) g[i] = (float)(sg * (double)g[i] * (double)a.scale);
        return;
    }
    // c_i = cumsum(y_true - y_pred)_i;  dL/dy_pred_j = - sum_{i>=j} f'(c_i)
    double c = 0.0, loss = 0.0;
    for (int i = 0; i < M; i++) {
        c += (double)yt[i] - (double)yp[i];
        loss += (a.kind == 0) ? fabs(c) : c * c;
        g[i] = (float)c;           // parked for the reverse pass
    }
    a.loss[r] = (float)((a.kind == 0) ? loss / M : loss);
    double tail = 0.0;
    for (int i = M - 1; i >= 0; i--) {
        const double ci = (double)g[i];
        tail += (a.kind == 0) ? ((ci > 0.0) ? 1.0 : (ci < 0.0 ? -1.0 : 0.0)) / M : 2.0 * ci;
        g[i] = (float)(-tail * (double)a.scale);
    }
}
