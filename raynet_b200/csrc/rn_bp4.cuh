// rn_bp4.cuh -- one BP sweep (a5 + a6) on the resident layout: register-resident rays, rows staged by TMA.
//
// History of this kernel (numbers: B200, C3, ms per non-first sweep; profiles/README.md):
//   bp2_kernel  rows staged by TMA bulk copies, per-voxel state in shared memory        5.25
//   (bp3)       state in registers per length class, rows loaded straight to registers  4.70
//   bp4 round 1 bp3 + the NEXT ray's rows prefetched into shared memory by cp.async     3.73
//   bp4 round 2 rows staged by TMA (one elected lane, three cp.async.bulk per ray, mbarrier)   4.15
//               the same with the per-lane cp.async staging                               3.78
// ncu on the round-1 version: L1TEX data pipe 75 % busy, 3 * NCH LDGSTS per lane per ray for the row
// staging on top of the gathers, the REDs and the transposition traffic.  Round 2 built the TMA staging
// the north star asks for (RN_BP4_TMA=1: UBLKCP, no LSU / L1TEX involvement, completion on an mbarrier,
// proxy fence before a buffer is handed back to the async proxy) and measured it SLOWER than the per-lane
// copies -- rows of ~1.3 KB are too small for one bulk operation each -- so it is a build option, not the
// default.  The kernel is instantiated for every length class up to RN_MAX_NCH = 12 chunks (1536 voxels:
// C5), with CTAs of two warps for the long classes so that the double-buffered rows of a CTA still allow
// several CTAs per SM; the transposition between the lane-consecutive gather / RED layout and the
// 4-voxels-per-lane scan layout goes through the s_hat chunk that has just been consumed.
//
// Rays are binned by length class (rn_class_of: NCH = ceil(L / 128) chunks); the kernel is instantiated
// per class and fully unrolled: every per-voxel quantity that has to survive from the forward to the
// backward pass (w_i, cp_i s_i, prefix sums) lives in registers.
// Contract: every ray of the launch has exactly NCH chunks (guaranteed by the binning) -- only the
// last chunk is masked.
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

#ifndef RN_BP4_ABLATE
#define RN_BP4_ABLATE 0        // 1 / 2: timing experiments without the REDs / the gathers (results are wrong)
#endif
#ifndef RN_BP4_RAYS_PER_WARP
#define RN_BP4_RAYS_PER_WARP 8
#endif
#ifndef RN_BP4_WAIT_ONE
#define RN_BP4_WAIT_ONE 0
#endif
#ifndef RN_BP4_TMA_HINT
#define RN_BP4_TMA_HINT 1      // L2 evict-first hint on the bulk copies
#endif
#ifndef RN_BP4_TMA_FENCE
#define RN_BP4_TMA_FENCE 1     // proxy fence before a buffer goes back to the async proxy
#endif
#ifndef RN_BP4_TMA
// 0: rows staged by per-lane 16-byte cp.async (LDGSTS); 1: by cp.async.bulk + mbarrier (TMA, UBLKCP).
// Measured on C3 (B200, same box, ms per non-first sweep; profiles/README.md): LDGSTS 3.78, TMA 4.15 -- a ray's
// three rows are ~1.3 KB each, one bulk copy per row per ray (12 small bulk operations per microsecond and SM)
// costs more than the LSU slots the per-lane copies take, so the per-lane copies are the default.
#define RN_BP4_TMA 0
#endif

__device__ __forceinline__ void rn_bulk_g2s_nohint(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void rn_cp_async16(uint32_t dst_smem, const void *src, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void rn_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void rn_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// warps per CTA of a length class: the double-buffered rows of one warp take 24 * NCH * 128 bytes
#ifndef RN_BP4_W4_MAX
#define RN_BP4_W4_MAX 6       // length classes up to this many chunks use CTAs of 4 warps, longer ones of 2
#endif
#ifndef RN_BP4_W8_MAX
#define RN_BP4_W8_MAX 0       // ... and up to this many chunks CTAs of 8 warps
#endif
__host__ __device__ constexpr int rn_bp4_warps(int nch) { return nch <= RN_BP4_W8_MAX ? 8 : (nch <= RN_BP4_W4_MAX ? 4 : 2); }
// shared memory words of one warp: 2 x (lin, s_hat) [+ 2 x msgs unless first sweep] rows of NCH * 128 words.
// The transposition scratch of chunk c is the 128 words of the CURRENT ray's s_hat chunk c, dead once its
// four values per lane have been read.
__host__ __device__ constexpr int rn_bp4_warp_words(int nch, bool first) { return nch * RN_CHUNK * (first ? 4 : 6); }

// One ray of a sweep, shared by bp4_kernel and bp4_first_mapped_kernel.  sLin / sS (/ sM unless kFirst): the ray's
// rows in shared memory, NCH chunks of 128; sS doubles as the transposition scratch (chunk c is dead once its four
// values per lane have been read).  uniform (first sweep straight after a reset): the accumulator is the prior
// everywhere -- one load, no gathers.
template <int NCH, bool kFirst>
__device__ __forceinline__ void rn_bp4_ray(const float *acc_in, float *acc_out, const bool uniform, const int *sLin, float *sS,
                                           const float *sM, float *m_row, const int L, const int lane,
                                           const uint64_t pol_stream, const uint64_t pol_keep) {
    // ---- accumulator gathers, lane-consecutive (neighbouring lanes share sectors) -------------
    // (first sweep straight after a reset: the accumulator is the prior everywhere -- one load)
    float ga[NCH][4];
    if (uniform) {
        const float acc0 = __ldg(acc_in);
#pragma unroll
        for (int c = 0; c < NCH; c++)
#pragma unroll
            for (int j = 0; j < 4; j++) ga[c][j] = acc0;
    } else {
#pragma unroll
        for (int c = 0; c < NCH; c++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = c * RN_CHUNK + 32 * j + lane;
                ga[c][j] = 0.f;
#if RN_BP4_ABLATE == 2   // timing experiment only: no accumulator gathers
                if (c < NCH - 1 || i < L) ga[c][j] = (float)sLin[i] * 1e-12f;
#else
                if (c < NCH - 1 || i < L) ga[c][j] = rn_ld_acc_pol(acc_in + sLin[i], pol_keep);
#endif
            }
        }
    }

    // ---- forward: occupancy-to-ray values, prefix scans ----------------------------------------
    float w[NCH][4], cps[NCH][4], pre0[NCH], tot[NCH];
    float carry_cp = 1.f, carry_pre = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        const int i0 = c * RN_CHUNK + 4 * lane;
        const float4 s4 = *reinterpret_cast<const float4 *>(sS + i0);
        float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f);   // first sweep: messages are 0 (mrf_np.py:275)
        if (!kFirst) m4 = *reinterpret_cast<const float4 *>(sM + i0);
        float accv[4];
        if (uniform) {
#pragma unroll
            for (int j = 0; j < 4; j++) accv[j] = ga[c][j];
        } else {   // lane-consecutive -> 4 consecutive voxels per lane, through the s_hat chunk just read
            float *sX = sS + c * RN_CHUNK;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; j++) sX[32 * j + lane] = ga[c][j];
            __syncwarp();
            const float4 acc4 = *reinterpret_cast<const float4 *>(sX + 4 * lane);
            accv[0] = acc4.x; accv[1] = acc4.y; accv[2] = acc4.z; accv[3] = acc4.w;
        }
        float mv[4] = {m4.x, m4.y, m4.z, m4.w};
        float sv[4] = {s4.x, s4.y, s4.z, s4.w};
        float o[4], q[4];
        if (c == NCH - 1) {
            // slots beyond the ray: s = 0 (nothing reaches the sums), message 0 (accumulator is 0 already)
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
        // exclusive products cp_i = prod_{k<i} (1 - o_k)
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
        // prefix sums of a_i = o_i cp_i s_i (true exclusive scan: no cancellation)
        const float la = fmaf(o[3], cps[c][3], fmaf(o[2], cps[c][2], fmaf(o[1], cps[c][1], o[0] * cps[c][0])));
        const float sinc = rn_warp_incl_scan_add(la, lane);
        float sexc = __shfl_up_sync(RN_FULL_MASK, sinc, 1);
        if (lane == 0) sexc = 0.f;
        tot[c] = __shfl_sync(RN_FULL_MASK, sinc, 31);
        pre0[c] = carry_pre + sexc;
        carry_pre += tot[c];
    }

    // ---- backward: suffix sums, messages, scatter-add ----------------------------------------------
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
            // p / (1 - p) = pos / neg, pos = pre + cp s, neg = pre + suf / q  ->  pos q / (pre q + suf)
            const float pos = pre + cps[c][j];
            const float den = fmaf(pre, q[j], suf[j]);
            msg[j] = 0.6931471805599453f * rn_lg2((pos * q[j]) * rn_rcp(den));
            pre += av[j];
        }
        carry_suf += tot[c];
        const int i0 = c * RN_CHUNK + 4 * lane;
        const float4 msg4 = make_float4(msg[0], msg[1], msg[2], msg[3]);
        if (c < NCH - 1 || i0 < L) rn_st_stream4_pol(m_row + i0, msg4, pol_stream);   // rows hold whole quads
        float *sX = sS + c * RN_CHUNK;   // every chunk has its own scratch: no wait for the previous chunk's readers
        *reinterpret_cast<float4 *>(sX + 4 * lane) = msg4;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = c * RN_CHUNK + 32 * j + lane;
#if RN_BP4_ABLATE == 1   // timing experiment only: no scatter-adds (the value still has to be produced)
            if ((c < NCH - 1 || i < L) && sX[32 * j + lane] == 1.2345e-30f) rn_red_add_pol(acc_out + sLin[i], sX[32 * j + lane], pol_keep);
#else
            if (c < NCH - 1 || i < L) rn_red_add_pol(acc_out + sLin[i], sX[32 * j + lane], pol_keep);
#endif
        }
    }
}

template <int NCH, bool kFirst>
__global__ void __launch_bounds__(32 * rn_bp4_warps(NCH)) bp4_kernel(RnDev p, Bp2Args a) {
    extern __shared__ __align__(128) unsigned char rn_bp4_smem[];
    constexpr int ROW = NCH * RN_CHUNK;
    constexpr int WARPS = rn_bp4_warps(NCH);
    constexpr bool kTma = RN_BP4_TMA != 0;
    __shared__ __align__(8) uint64_t bars[WARPS][2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float *base = reinterpret_cast<float *>(rn_bp4_smem) + (size_t)wid * rn_bp4_warp_words(NCH, kFirst);
    // layout: lin[2][ROW], s_hat[2][ROW], (msgs[2][ROW])
    const uint64_t pol_stream = rn_policy_evict_first();
    const uint64_t pol_keep = rn_policy_evict_last();

    if (kTma) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int w = 0; w < WARPS; w++) {
                rn_mbar_init(rn_smem_u32(&bars[w][0]), 1);
                rn_mbar_init(rn_smem_u32(&bars[w][1]), 1);
            }
            rn_mbar_init_fence();
        }
        __syncthreads();
    }

    // this warp's rays: positions k0 + WARPS * t of the launch (the warps of the CTA hold consecutive
    // entries of order[], i.e. neighbouring pixels, at any time)
    const int rpw = a.rays_per_warp;
    const int64_t k0 = (int64_t)blockIdx.x * (WARPS * rpw) + wid;
    int nmine = 0;
    if (k0 < a.n) nmine = (int)min((int64_t)rpw, (a.n - k0 + WARPS - 1) / WARPS);
    if (nmine == 0) return;

    auto ray_of = [&](int t) -> int64_t {
        const int64_t k = k0 + WARPS * (int64_t)t;
        return a.order ? (int64_t)__ldg(a.order + a.first + k) : a.first + k;
    };
    auto prefetch = [&](int64_t r, int L, int b) {
        const int32_t *lin_row = a.lin + r * (int64_t)p.row_stride;
        const float *s_row = a.s_hat + r * (int64_t)p.row_stride;
        const float *m_row = a.msgs + r * (int64_t)p.row_stride;
        if (kTma) {
            if (lane == 0) {   // rows hold whole quads: 16-byte multiples from 512-byte aligned rows
                const uint32_t bytes = (uint32_t)((L + 3) & ~3) * 4u;
                const uint32_t bar = rn_smem_u32(&bars[wid][b]);
                rn_mbar_expect_tx(bar, bytes * (kFirst ? 2u : 3u));
#if RN_BP4_TMA_HINT
                rn_bulk_g2s(rn_smem_u32(base + b * ROW), lin_row, bytes, bar, pol_stream);
                rn_bulk_g2s(rn_smem_u32(base + (2 + b) * ROW), s_row, bytes, bar, pol_stream);
                if (!kFirst) rn_bulk_g2s(rn_smem_u32(base + (4 + b) * ROW), m_row, bytes, bar, pol_stream);
#else
                rn_bulk_g2s_nohint(rn_smem_u32(base + b * ROW), lin_row, bytes, bar);
                rn_bulk_g2s_nohint(rn_smem_u32(base + (2 + b) * ROW), s_row, bytes, bar);
                if (!kFirst) rn_bulk_g2s_nohint(rn_smem_u32(base + (4 + b) * ROW), m_row, bytes, bar);
#endif
            }
        } else {
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const int i0 = c * RN_CHUNK + 4 * lane;
                if (c < NCH - 1 || i0 < L) {
                    rn_cp_async16(rn_smem_u32(base + b * ROW + i0), lin_row + i0, pol_stream);
                    rn_cp_async16(rn_smem_u32(base + (2 + b) * ROW + i0), s_row + i0, pol_stream);
                    if (!kFirst) rn_cp_async16(rn_smem_u32(base + (4 + b) * ROW + i0), m_row + i0, pol_stream);
                }
            }
            rn_cp_async_commit();
        }
    };

    int64_t r_cur = ray_of(0);
    int L_cur = __ldg(a.count + r_cur);
    prefetch(r_cur, L_cur, 0);
    int64_t r_nxt = 0;
    int L_nxt = 0;
    if (nmine > 1) { r_nxt = ray_of(1); L_nxt = __ldg(a.count + r_nxt); }

    for (int t = 0; t < nmine; t++) {
        const int b = t & 1;
        const int64_t r = r_cur;
        const int L = L_cur;
        // every lane is done with the buffers of ray t - 1; its generic-proxy stores into them (transposition
        // scratch) are ordered before the async-proxy writes of the next bulk copy
        if (kTma && RN_BP4_TMA_FENCE) rn_fence_async_smem();
        __syncwarp();
        if (t + 1 < nmine) {
            prefetch(r_nxt, L_nxt, b ^ 1);
            r_cur = r_nxt; L_cur = L_nxt;
            if (t + 2 < nmine) { r_nxt = ray_of(t + 2); L_nxt = __ldg(a.count + r_nxt); }
            if (!kTma) rn_cp_async_wait<1>();
        } else if (!kTma) {
            rn_cp_async_wait<0>();
        }
        if (kTma) {
#if RN_BP4_WAIT_ONE
            if (lane == 0) rn_mbar_wait(rn_smem_u32(&bars[wid][b]), (uint32_t)((t >> 1) & 1));
            __syncwarp();
#else
            rn_mbar_wait(rn_smem_u32(&bars[wid][b]), (uint32_t)((t >> 1) & 1));   // buffer b: use number t / 2
#endif
        } else {
            __syncwarp();   // ... and every lane's copies of ray t have landed
        }
        const int *sLin = reinterpret_cast<const int *>(base + b * ROW);
        float *sS = base + (2 + b) * ROW;
        const float *sM = base + (4 + b) * ROW;   // !kFirst only
        float *m_row = a.msgs + r * (int64_t)p.row_stride;

        rn_bp4_ray<NCH, kFirst>(a.acc_in, a.acc_out, kFirst && a.uniform_acc, sLin, sS, sM, m_row, L, lane, pol_stream, pol_keep);
    }
}


// =======================================================================================
// a8 + a9 on the resident layout: depth re-estimation + arg-max -> depth, warp per ray.
// depth2_kernel (rn_engine.cuh) serves every layout and the S_new output of the tests and was
// issue-bound (81 % of the issue slots, 1070 warp instructions per ray); this is its resident-only
// fast path: masks in the last chunk only, cache policies created once, o / q formed directly
// from min(o, 1 - o), the reference image of a ray looked up by one lane with a guess-and-check.
// =======================================================================================
__device__ __forceinline__ void rn_occ_oq(float acc, float msg, float &o, float &q) {
    const float x = acc - msg;
    const float u = fmaxf(rn_rcp(1.0f + rn_ex2(fabsf(x) * 1.4426950408889634f)), 1e-4f);   // min(o, 1 - o), clipped
    const bool pos = x >= 0.f;
    o = pos ? 1.0f - u : u;
    q = pos ? u : 1.0f - u;
}

// One chunk (128 voxels) of a ray: the row quads, the gathered accumulator values of the chunk and the accumulator
// offsets of the NEXT chunk, so that its gathers do not wait for an offset load.
struct Depth3Stage {
    float ga[4];
    float4 s4, m4;
    int32_t lin_next[4];
};

// loads of chunk c: gathers through `lin` (fetched one chunk earlier), the s / msgs quads, and the offsets of chunk c + 1
__device__ __forceinline__ void rn_depth3_load(const Depth2Args &a, const int32_t *lin_row, const float *s_row, const float *m_row,
                                               int c, int L, int lane, uint64_t pol_stream, uint64_t pol_keep,
                                               const int32_t (&lin)[4], Depth3Stage &st) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = c * RN_CHUNK + 32 * j + lane;
        st.ga[j] = (i < L) ? rn_ld_acc_pol(a.acc + lin[j], pol_keep) : 0.f;
        st.lin_next[j] = (i + RN_CHUNK < L) ? rn_ld_stream_s32(lin_row + i + RN_CHUNK, pol_stream) : 0;
    }
    const int i0 = c * RN_CHUNK + 4 * lane;
    st.s4 = st.m4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i0 < L) {
        st.s4 = rn_ld_stream4_pol(s_row + i0, pol_stream);
        st.m4 = rn_ld_stream4_pol(m_row + i0, pol_stream);
    }
}

template <bool kTail>
__device__ __forceinline__ void rn_depth3_chunk(const Depth3Stage &st, float *sX, int c, int L, int lane,
                                                float &carry_cp, float &bestv, int &besti) {
    const int i0 = c * RN_CHUNK + 4 * lane;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; j++) sX[32 * j + lane] = st.ga[j];
    __syncwarp();
    const float4 acc4 = *reinterpret_cast<const float4 *>(sX + 4 * lane);
    const float accv[4] = {acc4.x, acc4.y, acc4.z, acc4.w};
    float mv[4] = {st.m4.x, st.m4.y, st.m4.z, st.m4.w};
    float sv[4] = {st.s4.x, st.s4.y, st.s4.z, st.s4.w};
    float o[4], q[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (kTail) {   // slots beyond the ray: s = 0 -> a_i = 0, never the maximum of a ray with L > 1
            const bool ok = i0 + j < L;
            sv[j] = ok ? sv[j] : 0.f;
            mv[j] = ok ? mv[j] : 0.f;
        }
        rn_occ_oq(accv[j], mv[j], o[j], q[j]);
    }
    const float lp0 = q[0], lp1 = lp0 * q[1], lp2 = lp1 * q[2], lp3 = lp2 * q[3];
    const float inc = rn_warp_incl_scan_mul(lp3, lane);
    float exc = __shfl_up_sync(RN_FULL_MASK, inc, 1);
    if (lane == 0) exc = 1.f;
    const float base = carry_cp * exc;
    carry_cp = carry_cp * __shfl_sync(RN_FULL_MASK, inc, 31);
    const float av[4] = {o[0] * (base * sv[0]), o[1] * ((base * lp0) * sv[1]), o[2] * ((base * lp1) * sv[2]),
                         o[3] * ((base * lp2) * sv[3])};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool better = kTail ? (i0 + j < L && av[j] > bestv) : (av[j] > bestv);
        bestv = better ? av[j] : bestv;
        besti = better ? i0 + j : besti;
    }
}

__global__ void __launch_bounds__(128) depth3_kernel(RnDev p, Depth2Args a) {
    __shared__ __align__(16) float sXall[4][128];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * 4 + wid;
    if (r >= a.n_rays) return;
    float *sX = sXall[wid];
    const int L = __ldg(a.count + r);
    const uint64_t pol_stream = rn_policy_evict_first();
    const uint64_t pol_keep = rn_policy_evict_last();
    const int32_t *lin_row = a.lin + r * (int64_t)p.row_stride;
    const float *s_row = a.s_hat + r * (int64_t)p.row_stride;
    const float *m_row = a.msgs + r * (int64_t)p.row_stride;
    float bestv = -INFINITY;
    int besti = 0;
    if (L > 1) {   // mrf_np.py:376-377: rays with count <= 1 keep an all-zero row
        const int nch = (L + RN_CHUNK - 1) / RN_CHUNK;
        float carry_cp = 1.f;
        // the accumulator offsets run one chunk ahead of the gathers that use them
        int32_t lin0[4];
#pragma unroll
        for (int j = 0; j < 4; j++) lin0[j] = (32 * j + lane < L) ? rn_ld_stream_s32(lin_row + 32 * j + lane, pol_stream) : 0;
        Depth3Stage cur, nxt;
        rn_depth3_load(a, lin_row, s_row, m_row, 0, L, lane, pol_stream, pol_keep, lin0, cur);
        for (int c = 0; c < nch - 1; c++) {
            // (issuing these loads BEFORE the scan of chunk c -- two chunks in flight, 56 registers -- measured 2.56 ms
            // on C3 against 2.27 ms in this order; without the early offsets 2.39-2.45 ms)
            rn_depth3_chunk<false>(cur, sX, c, L, lane, carry_cp, bestv, besti);
            rn_depth3_load(a, lin_row, s_row, m_row, c + 1, L, lane, pol_stream, pol_keep, cur.lin_next, nxt);
            cur = nxt;
        }
        rn_depth3_chunk<true>(cur, sX, nch - 1, L, lane, carry_cp, bestv, besti);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {   // first maximum over the ray (raynet_fp.py:193-205)
            const float ov = __shfl_xor_sync(RN_FULL_MASK, bestv, d);
            const int oi = __shfl_xor_sync(RN_FULL_MASK, besti, d);
            if (ov > bestv || (ov == bestv && oi < besti)) { bestv = ov; besti = oi; }
        }
    }
    if (lane == 0) {
        // voxel of the arg-max slot; an all-zero row selects slot 0 (the first voxel of the ray or, for an
        // empty ray, the zero-filled triplet (0, 0, 0); raynet_fp.py:206-226)
        int x = 0, y = 0, z = 0;
        if (L >= 1) rn_unbrick(p, __ldg(lin_row + ((L > 1) ? besti : 0)), x, y, z);
        int seg = 0;
        if (a.seg_starts) {
            // usual case: every image has the same number of rays -> the image is r / rays-per-image;
            // checked against the table, binary search otherwise
            const int64_t len = __ldg(a.seg_starts + 1);
            const int guess = len > 0 ? (int)min((int64_t)(a.n_seg - 1), r / len) : 0;
            if (__ldg(a.seg_starts + guess) <= r && r < __ldg(a.seg_starts + guess + 1)) {
                seg = guess;
            } else {
                int lo_s = 0, hi_s = a.n_seg;
                while (hi_s - lo_s > 1) {
                    const int mid = (lo_s + hi_s) >> 1;
                    if (__ldg(a.seg_starts + mid) <= r) lo_s = mid; else hi_s = mid;
                }
                seg = lo_s;
            }
        }
        const float *C = a.centres + 4 * seg;
        const float cc[3] = {__ldg(a.axes + x), __ldg(a.axes + p.gx + y), __ldg(a.axes + p.gx + p.gy + z)};
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 3; i++) { const float dd = cc[i] - __ldg(C + i); sum += dd * dd; }
        a.depth_map[r] = sqrtf(sum);
    }
}
