// rn_bp3.cuh -- one BP sweep (a5 + a6) with the whole ray held in REGISTERS.
//
// Why a third version.  ncu on bp2_kernel (profiles/r01_*) shows the L1TEX data pipe as the
// limiter: 69 % busy, two thirds of its wavefronts being SHARED-memory traffic (TMA row
// staging, LDS/STS of the per-voxel state between the forward and the backward pass), the rest
// the accumulator gathers and the RED scatter-adds, which by scratch/mb_l1.cu cost ~1 cycle per
// 128-byte line (LDG) and ~1.5 cycles per 32-byte sector (RED) and cannot be avoided.  Rays are
// binned by length class (rn_class_of: NCH = ceil(L / 128) chunks), so the kernel is
// instantiated per class and fully unrolled: every per-voxel quantity that has to survive from
// the forward to the backward pass (w_i, cp_i s_i, prefix sums, accumulator offsets) lives in
// registers, the s_hat / message rows are read and written with coalesced 128-bit global
// accesses directly in the "4 consecutive voxels per lane" layout the scans want, and shared
// memory is used for one thing only: turning the lane-consecutive gather / RED layout (voxel
// 32 j + lane: neighbouring lanes share sectors) into that layout and back (4 + 4 wavefronts per
// direction per 128 voxels instead of ~95 in bp2).
//
// Arithmetic: identical to bp2_kernel (see there for the cancellation-free forms) except that
// pos / neg is evaluated as pos q / (pre q + suf): one MUFU.RCP less per voxel.
#pragma once

#include "rn_engine.cuh"

__device__ __forceinline__ int rn_ld_stream_s32(const int32_t *p, uint64_t pol) {
    int v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 rn_ld_stream4_pol(const float *p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void rn_st_stream4_pol(float *p, float4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
                 : "memory");
}

// min(o, 1 - o) = 1 / (1 + e^{|x|}) for o = sigmoid(x), clipped at 1e-4 (mrf_np.py:52-71); signed
// like rn_occ_w.  e^{|x|} may overflow to +inf: rcp(inf) = 0 -> clipped to 1e-4, as intended.
__device__ __forceinline__ float rn_occ_w2(float acc, float msg) {
    const float x = acc - msg;
    const float u = fmaxf(rn_rcp(1.0f + rn_ex2(fabsf(x) * 1.4426950408889634f)), 1e-4f);
    return (x >= 0.f) ? -u : u;
}

#ifndef RN_BP3_RAYS_PER_CTA
#define RN_BP3_RAYS_PER_CTA 64
#endif

template <int NCH, bool kFirst>
__global__ void __launch_bounds__(128) bp3_kernel(RnDev p, Bp2Args a) {
    __shared__ __align__(16) float sXall[4][NCH][RN_CHUNK];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t pol_stream = rn_policy_evict_first();
    const uint64_t pol_keep = rn_policy_evict_last();
    // A CTA works through RN_BP3_RAYS_PER_CTA consecutive entries of order[] (neighbouring pixels of
    // one 8x8 tile for the most part): its four warps always hold four adjacent rays, and the
    // sectors of the accumulator that the first rays pulled into L1 serve the gathers of the others.
    for (int it = 0; it < RN_BP3_RAYS_PER_CTA / 4; it++) {
    const int64_t k = (int64_t)blockIdx.x * RN_BP3_RAYS_PER_CTA + it * 4 + wid;
    if (k >= a.n) break;
    const int64_t r = a.order ? (int64_t)__ldg(a.order + a.first + k) : a.first + k;
    const int L = __ldg(a.count + r);
    if (L <= 1) continue;   // mrf_np.py:299-301
    __syncwarp();           // the previous ray's reads of the transposition scratch are done
    const int32_t *lin_row = a.lin + r * (int64_t)p.row_stride;
    const float *s_row = a.s_hat + r * (int64_t)p.row_stride;
    float *m_row = a.msgs + r * (int64_t)p.row_stride;

    // ---- every load of the ray is issued before anything is consumed ---------------------------
    int lin[NCH][4];
    float ga[NCH][4];
    float4 s4[NCH], m4[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = c * RN_CHUNK + 32 * j + lane;
            lin[c][j] = (i < L) ? rn_ld_stream_s32(lin_row + i, pol_stream) : -1;
        }
    }
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        const int i0 = c * RN_CHUNK + 4 * lane;
        s4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        m4[c] = s4[c];
        if (i0 < L) {
            s4[c] = rn_ld_stream4_pol(s_row + i0, pol_stream);
            if (!kFirst) m4[c] = rn_ld_stream4_pol(m_row + i0, pol_stream);
        }
    }
#pragma unroll
    for (int c = 0; c < NCH; c++) {
#pragma unroll
        for (int j = 0; j < 4; j++) ga[c][j] = (lin[c][j] >= 0 && !(a.debug & 2)) ? rn_ld_acc_pol(a.acc_in + lin[c][j], pol_keep) : 0.f;
    }

    // ---- forward: occupancy-to-ray values, prefix scans ------------------------------------------
    float w[NCH][4], cps[NCH][4], pre0[NCH], tot[NCH];
    float carry_cp = 1.f, carry_pre = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        float *sX = sXall[wid][c];
#pragma unroll
        for (int j = 0; j < 4; j++) sX[32 * j + lane] = ga[c][j];
        __syncwarp();
        const float4 acc4 = *reinterpret_cast<const float4 *>(sX + 4 * lane);
        const int i0 = c * RN_CHUNK + 4 * lane;
        const float accv[4] = {acc4.x, acc4.y, acc4.z, acc4.w};
        const float mv[4] = {m4[c].x, m4[c].y, m4[c].z, m4[c].w};
        const float sraw[4] = {s4[c].x, s4[c].y, s4[c].z, s4[c].w};
        float sv[4], o[4], q[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            // slots beyond the ray: s = 0 (nothing reaches the sums), message and accumulator 0
            const bool ok = i0 + j < L;
            sv[j] = ok ? sraw[j] : 0.f;
            w[c][j] = rn_occ_w2(accv[j], ok ? mv[j] : 0.f);
            rn_occ_from_w(w[c][j], o[j], q[j]);
        }
        // exclusive products cp_i = prod_{k<i} (1 - o_k)
        const float lp0 = q[0], lp1 = lp0 * q[1], lp2 = lp1 * q[2], lp3 = lp2 * q[3];
        const float inc = rn_warp_incl_scan_mul(lp3, lane);
        float exc = __shfl_up_sync(RN_FULL_MASK, inc, 1);
        if (lane == 0) exc = 1.f;
        const float base = carry_cp * exc;
        carry_cp = carry_cp * __shfl_sync(RN_FULL_MASK, inc, 31);
        cps[c][0] = base * sv[0];
        cps[c][1] = (base * lp0) * sv[1];
        cps[c][2] = (base * lp1) * sv[2];
        cps[c][3] = (base * lp2) * sv[3];
        // prefix sums of a_i = o_i cp_i s_i (true exclusive scan: no cancellation)
        const float la = fmaf(o[3], cps[c][3], fmaf(o[2], cps[c][2], fmaf(o[1], cps[c][1], o[0] * cps[c][0])));
        const float sinc = rn_warp_incl_scan_add(la, lane);
        float sexc = __shfl_up_sync(RN_FULL_MASK, sinc, 1);
        if (lane == 0) sexc = 0.f;
        tot[c] = __shfl_sync(RN_FULL_MASK, sinc, 31);
        pre0[c] = carry_pre + sexc;
        carry_pre += tot[c];
    }

    // ---- backward: suffix sums, messages, scatter-add ------------------------------------------------
    float carry_suf = 0.f;
#pragma unroll
    for (int c = NCH - 1; c >= 0; c--) {
        float o[4], q[4], av[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            rn_occ_from_w(w[c][j], o[j], q[j]);
            av[j] = o[j] * cps[c][j];
        }
        const float ra3 = av[3], ra2 = av[2] + ra3, ra1 = av[1] + ra2, ra0 = av[0] + ra1;
        // sum over the lanes ABOVE this one: shift, then inclusive reverse scan (exact exclusive)
        float above = __shfl_down_sync(RN_FULL_MASK, ra0, 1);
        if (lane == 31) above = 0.f;
        const float sbase = carry_suf + rn_warp_incl_rscan_add(above, lane);
        const float suf[4] = {sbase + ra1, sbase + ra2, sbase + ra3, sbase};
        float pre = pre0[c];
        float msg[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            // p / (1 - p) = pos / neg with pos = pre + cp s, neg = pre + suf / q  ->  pos q / (pre q + suf)
            const float pos = pre + cps[c][j];
            const float den = fmaf(pre, q[j], suf[j]);
            msg[j] = 0.6931471805599453f * rn_lg2((pos * q[j]) * rn_rcp(den));
            pre += av[j];
        }
        carry_suf += tot[c];
        const int i0 = c * RN_CHUNK + 4 * lane;
        const float4 msg4 = make_float4(msg[0], msg[1], msg[2], msg[3]);
        if (i0 < L && !(a.debug & 4)) rn_st_stream4_pol(m_row + i0, msg4, pol_stream);   // rows hold whole quads (row_stride % 128 == 0)
        float *sX = sXall[wid][c];
        *reinterpret_cast<float4 *>(sX + 4 * lane) = msg4;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (lin[c][j] >= 0 && !(a.debug & 1)) rn_red_add_pol(a.acc_out + lin[c][j], sX[32 * j + lane], pol_keep);
    }
    }
}

// Diagnostics only (RN_BP_DEBUG=8): the memory operations of one sweep with no arithmetic in between --
// the floor the memory system sets for this access pattern.
template <int NCH>
__global__ void __launch_bounds__(128) bp_memonly_kernel(RnDev p, Bp2Args a) {
    __shared__ __align__(16) float sXall[4][RN_CHUNK];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t pol_stream = rn_policy_evict_first();
    const uint64_t pol_keep = rn_policy_evict_last();
    const int64_t k = (int64_t)blockIdx.x * 4 + wid;
    if (k >= a.n) return;
    const int64_t r = a.order ? (int64_t)__ldg(a.order + a.first + k) : a.first + k;
    const int L = __ldg(a.count + r);
    const int32_t *lin_row = a.lin + r * (int64_t)p.row_stride;
    const float *s_row = a.s_hat + r * (int64_t)p.row_stride;
    float *m_row = a.msgs + r * (int64_t)p.row_stride;
    float *sX = sXall[wid];
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        int lin[4];
        float ga[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = c * RN_CHUNK + 32 * j + lane;
            lin[j] = (i < L) ? ((a.debug & 16) ? (int)((r * 131 + i) & 0xffffff) : rn_ld_stream_s32(lin_row + i, pol_stream)) : -1;
        }
        const int i0 = c * RN_CHUNK + 4 * lane;
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), m4 = s4;
        if (i0 < L) {
            s4 = rn_ld_stream4_pol(s_row + i0, pol_stream);
            m4 = rn_ld_stream4_pol(m_row + i0, pol_stream);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) ga[j] = (lin[j] >= 0 && !(a.debug & 2)) ? rn_ld_acc_pol(a.acc_in + lin[j], pol_keep) : 0.f;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; j++) sX[32 * j + lane] = ga[j];
        __syncwarp();
        const float4 acc4 = *reinterpret_cast<const float4 *>(sX + 4 * lane);
        const float4 msg4 = make_float4(s4.x + m4.x + acc4.x, s4.y + m4.y + acc4.y, s4.z + m4.z + acc4.z, s4.w + m4.w + acc4.w);
        if (i0 < L && !(a.debug & 4)) rn_st_stream4_pol(m_row + i0, msg4, pol_stream);
        __syncwarp();
        *reinterpret_cast<float4 *>(sX + 4 * lane) = msg4;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (lin[j] >= 0 && !(a.debug & 1)) rn_red_add_pol(a.acc_out + lin[j], sX[32 * j + lane], pol_keep);
    }
}
