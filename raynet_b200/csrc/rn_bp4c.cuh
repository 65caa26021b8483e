// rn_bp4c.cuh -- non-first BP sweep (a5 + a6) with the accumulator gathers and the scatter-adds of the WARPS rays of a
// CTA issued TOGETHER: one gather / RED instruction covers the same 32 / WARPS voxel slots of each of the CTA's rays
// instead of 32 slots of one ray.
//
// Why (profiles/README.md, "sector operations"): bp4_kernel is not bound by HBM but by 32-byte sector operations --
// every gather and every RED instruction of a ray touches ~16 sectors of the bricked accumulator (2 voxels per 2x2x2
// sector along a ray) and the machine retires ~1.9e11 RED sectors per second (profiles/r01_microbench_l1.txt), i.e.
// 2.1 ms of the 3.8 ms C3 sweep for the REDs alone.  The rays of a CTA are neighbouring pixels of one length class
// (consecutive entries of order[]): half a voxel apart, they run through the same bricks at the same slot numbers.
// scratch/sim_sectors.py on the C3 rig: 16.5 sectors per instruction for one ray, 7.9 for 8 slots of 4 vertical
// neighbours, 10.5 for 16 slots of 2.
//
// Everything else is bp4_kernel: one warp owns one ray, its per-voxel state lives in registers (rn_bp4.cuh), rows are
// staged one ray ahead by per-lane cp.async.  The redistribution between "slot s of ray r" (gather / RED) and "4
// consecutive slots per lane of the owner" (scans) goes through the owner's s_hat buffer like bp4's transposition, but
// across warps: three CTA barriers per ray instead of __syncwarp.  Each warp's buffers start 32 / WARPS words further
// into the bank cycle than its neighbour's, so the WARPS groups of an instruction hit disjoint banks.
#pragma once

#include "rn_bp4.cuh"

#ifndef RN_BP4C_MAP
#define RN_BP4C_MAP 1
#endif
__host__ __device__ constexpr int rn_bp4c_warp_words(int nch, int warps) { return nch * RN_CHUNK * 6 + 32 / warps; }

template <int NCH>
__global__ void __launch_bounds__(32 * rn_bp4_warps(NCH)) bp4c_kernel(RnDev p, Bp2Args a) {
    extern __shared__ __align__(128) unsigned char rn_bp4_smem[];
    constexpr int ROW = NCH * RN_CHUNK;
    constexpr int WARPS = rn_bp4_warps(NCH);
    constexpr int LPR = 32 / WARPS;                       // lanes per ray of a gather / RED instruction
    constexpr int WORDS = rn_bp4c_warp_words(NCH, WARPS);
    __shared__ int sL[2][WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float *smem = reinterpret_cast<float *>(rn_bp4_smem);
    float *base = smem + (size_t)wid * WORDS;             // lin[2][ROW], s_hat[2][ROW], msgs[2][ROW]
    const uint64_t pol_stream = rn_policy_evict_first();
    const uint64_t pol_keep = rn_policy_evict_last();

    // the CTA's rays: entries k0c + WARPS * t + w of the launch, t < n_iter; warp w owns the w-th of every group
    const int rpw = a.rays_per_warp;
    const int64_t k0c = (int64_t)blockIdx.x * (WARPS * rpw);
    if (k0c >= a.n) return;
    const int n_iter = (int)min((int64_t)rpw, (a.n - k0c + WARPS - 1) / WARPS);
    const int64_t k0 = k0c + wid;
    int nmine = 0;
    if (k0 < a.n) nmine = (int)min((int64_t)rpw, (a.n - k0 + WARPS - 1) / WARPS);

    auto ray_of = [&](int t) -> int64_t {
        const int64_t k = k0 + WARPS * (int64_t)t;
        return a.order ? (int64_t)__ldg(a.order + a.first + k) : a.first + k;
    };
    auto prefetch = [&](int64_t r, int L, int b) {
        const int32_t *lin_row = a.lin + r * (int64_t)p.row_stride;
        const float *s_row = a.s_hat + r * (int64_t)p.row_stride;
        const float *m_row = a.msgs + r * (int64_t)p.row_stride;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const int i0 = c * RN_CHUNK + 4 * lane;
            if (i0 < L) {
                rn_cp_async16(rn_smem_u32(base + b * ROW + i0), lin_row + i0, pol_stream);
                rn_cp_async16(rn_smem_u32(base + (2 + b) * ROW + i0), s_row + i0, pol_stream);
                rn_cp_async16(rn_smem_u32(base + (4 + b) * ROW + i0), m_row + i0, pol_stream);
            }
        }
        rn_cp_async_commit();
    };

    int64_t r_cur = 0, r_nxt = 0;
    int L_cur = 0, L_nxt = 0;
    if (nmine > 0) { r_cur = ray_of(0); L_cur = __ldg(a.count + r_cur); prefetch(r_cur, L_cur, 0); }
    if (nmine > 1) { r_nxt = ray_of(1); L_nxt = __ldg(a.count + r_nxt); }

    // the slots this lane gathers / scatter-adds: ray rr of the group, slots sl0 + 128 c + LPR (4 ... ) below
#if RN_BP4C_MAP == 0
    const int rr = lane / LPR;
    const int sl0 = wid * 4 * LPR + (lane % LPR);
#else   // neighbouring lanes = the same slot of neighbouring rays
    const int rr = lane % WARPS;
    const int sl0 = wid * 4 * LPR + (lane / WARPS);
#endif
    float *const grp = smem + (size_t)rr * WORDS;

    for (int t = 0; t < n_iter; t++) {
        const int b = t & 1;
        const bool mine = t < nmine;
        const int64_t r = r_cur;
        const int L = mine ? L_cur : 0;
        const int *sLin = reinterpret_cast<const int *>(base + b * ROW);
        float *sS = base + (2 + b) * ROW;
        const float *sM = base + (4 + b) * ROW;
        float *m_row = a.msgs + r * (int64_t)p.row_stride;

        // this thread's own copies have landed: the s_hat quads it owns go to registers, which frees the s_hat
        // buffer as the redistribution scratch of this ray
        rn_cp_async_wait<0>();
        float4 s4[NCH];
#pragma unroll
        for (int c = 0; c < NCH; c++) s4[c] = *reinterpret_cast<const float4 *>(sS + c * RN_CHUNK + 4 * lane);
        if (lane == 0) sL[b][wid] = L;
        __syncthreads();   // rows of every ray of the group visible; everybody is done with the previous group's buffers
        if (t + 1 < nmine) {
            prefetch(r_nxt, L_nxt, b ^ 1);
            r_cur = r_nxt; L_cur = L_nxt;
            if (t + 2 < nmine) { r_nxt = ray_of(t + 2); L_nxt = __ldg(a.count + r_nxt); }
        }

        // ---- accumulator gathers: LPR slots of each of the WARPS rays per instruction ---------------------
        const int Lrr = sL[b][rr];
        const int *gLin = reinterpret_cast<const int *>(grp + b * ROW);
        float *gX = grp + (2 + b) * ROW;
        {
            float ga[NCH][4];
#pragma unroll
            for (int c = 0; c < NCH; c++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int i = c * RN_CHUNK + j * LPR + sl0;
                    ga[c][j] = (i < Lrr) ? rn_ld_acc_pol(a.acc_in + gLin[i], pol_keep) : 0.f;
                }
#pragma unroll
            for (int c = 0; c < NCH; c++)
#pragma unroll
                for (int j = 0; j < 4; j++) gX[c * RN_CHUNK + j * LPR + sl0] = ga[c][j];
        }
        __syncthreads();   // gathered values of the own ray are in sS

        // ---- forward: occupancy-to-ray values, prefix scans (as rn_bp4_ray) ---------------------------------
        float w[NCH][4], cps[NCH][4], pre0[NCH], tot[NCH];
        float carry_cp = 1.f, carry_pre = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const int i0 = c * RN_CHUNK + 4 * lane;
            const float4 m4 = *reinterpret_cast<const float4 *>(sM + i0);
            const float4 acc4 = *reinterpret_cast<const float4 *>(sS + i0);
            const float accv[4] = {acc4.x, acc4.y, acc4.z, acc4.w};
            float mv[4] = {m4.x, m4.y, m4.z, m4.w};
            float sv[4] = {s4[c].x, s4[c].y, s4[c].z, s4[c].w};
            float o[4], q[4];
            if (c == NCH - 1) {   // slots beyond the ray: s = 0 (nothing reaches the sums), message 0
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const bool ok = i0 + j < L;
                    sv[j] = ok ? sv[j] : 0.f;
                    mv[j] = ok ? mv[j] : 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                w[c][j] = rn_occ_w2(accv[j], mv[j]);
                rn_occ_from_w(w[c][j], o[j], q[j]);
            }
            const float lp0 = q[0], lp1 = lp0 * q[1], lp2 = lp1 * q[2], lp3 = lp2 * q[3];
            const float inc = rn_warp_incl_scan_mul(lp3, lane);
            float exc = __shfl_up_sync(RN_FULL_MASK, inc, 1);
            if (lane == 0) exc = 1.f;
            const float basecp = carry_cp * exc;
            carry_cp = carry_cp * __shfl_sync(RN_FULL_MASK, inc, 31);
            cps[c][0] = basecp * sv[0];
            cps[c][1] = (basecp * lp0) * sv[1];
            cps[c][2] = (basecp * lp1) * sv[2];
            cps[c][3] = (basecp * lp2) * sv[3];
            const float la = fmaf(o[3], cps[c][3], fmaf(o[2], cps[c][2], fmaf(o[1], cps[c][1], o[0] * cps[c][0])));
            const float sinc = rn_warp_incl_scan_add(la, lane);
            float sexc = __shfl_up_sync(RN_FULL_MASK, sinc, 1);
            if (lane == 0) sexc = 0.f;
            tot[c] = __shfl_sync(RN_FULL_MASK, sinc, 31);
            pre0[c] = carry_pre + sexc;
            carry_pre += tot[c];
        }

        // ---- backward: suffix sums, messages --------------------------------------------------------------
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
            float above = __shfl_down_sync(RN_FULL_MASK, ra0, 1);
            if (lane == 31) above = 0.f;
            const float sbase = carry_suf + rn_warp_incl_rscan_add(above, lane);
            const float suf[4] = {sbase + ra1, sbase + ra2, sbase + ra3, sbase};
            float pre = pre0[c];
            float msg[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float pos = pre + cps[c][j];
                const float den = fmaf(pre, q[j], suf[j]);
                msg[j] = 0.6931471805599453f * rn_lg2((pos * q[j]) * rn_rcp(den));
                pre += av[j];
            }
            carry_suf += tot[c];
            const int i0 = c * RN_CHUNK + 4 * lane;
            const float4 msg4 = make_float4(msg[0], msg[1], msg[2], msg[3]);
            if (i0 < L) rn_st_stream4_pol(m_row + i0, msg4, pol_stream);   // rows hold whole quads
            *reinterpret_cast<float4 *>(sS + i0) = msg4;
        }
        __syncthreads();   // new messages of every ray of the group are in its sS

        // ---- scatter-add: the same slots of every ray of the group per instruction --------------------------
#pragma unroll
        for (int c = 0; c < NCH; c++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = c * RN_CHUNK + j * LPR + sl0;
                if (i < Lrr) rn_red_add_pol(a.acc_out + gLin[i], gX[i], pol_keep);
            }
    }
}
