/*
 * rn_oracle.c -- CPU restatement of the RayNet volumetric-inference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under raynet_b200/ may import, link or execute
 * this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / reported baseline.
 *
 * Every function follows the reference text it cites (paths relative to
 * /root/reference/raynet).  Arithmetic is written operation by operation in the
 * type the reference uses at that point (f32 vs f64), and the file is compiled with
 *   gcc -O2 -ffp-contract=off -fno-fast-math
 * so no FMA contraction happens: + - * / sqrt are IEEE-754 round-to-nearest in the
 * stated type.  The CUDA product code is compiled with -fmad=false and uses the same
 * operation order wherever an integer decision (voxel index, pixel index, bracket
 * index) depends on the value, which is what makes those decisions bit-exact.
 *
 * PINNING (see tests/test_oracle_pinning.py, DESIGN.md section "Oracle"):
 *   - voxel_traversal: reference golden vectors tests/test_ray_marching.py:20-77,92-102
 *     and bit-equality with the reference's own Cython build (oracle/_ref/ray_tracing*.so)
 *     on random rays.
 *   - bp / depth estimate / occupancy: the reference's mrf_np.py compiled from its own
 *     source (oracle/_ref/ref_mrf_np*.so) on the six tests/test_mrf.py scenarios and on
 *     random inputs; golden outputs committed under tests/golden/.
 *   - planes->voxels: reference numpy li / li_2 (oracle/_ref/ref_planes_voxels_mapping*.so),
 *     the cross-check of tests/test_planes_voxels_mapping.py:61-78.
 *   - sample_in_bbox, similarity, arg-max depth have NO CPU twin and no test in the
 *     reference: they are restated from the .cu text and pinned on the GPU box against an
 *     EXECUTION of the reference's own CUDA kernels (oracle/build_ref_cuda.py compiles
 *     the .cu files of cuda_implementations/ + the kernel text of raynet_fp.py for sm_100a,
 *     tests/test_gpu_ref_cuda.py runs them next to the sm_100a kernels and this oracle's
 *     outputs: voxel lists identical, distributions / messages / S_new within 1e-5).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define RN_EXPORT __attribute__((visibility("default")))

RN_EXPORT int rn_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

RN_EXPORT void rn_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline float clampf(float x, float a, float b) {
    /* cuda_implementations/utils.cu:1-3 : min(max(x, a), b) */
    return fminf(fmaxf(x, a), b);
}

/* ------------------------------------------------------------------------------------
 * a1. sample_in_bbox  -- cuda_implementations/sampling_schemes.cu:5-90
 *   pixel = (ray_idx / H, ray_idx % H)                                   (:5-8)
 *   X = P_inv(4x3) . (px, py, 1): products m*v in f32, accumulated in f64, the
 *       m*1.0 term is an f64 product, dehomogenised in f64               (:15-39)
 *   dir = (f64 X) - (f32 centre) stored f32                              (:58-60)
 *   slab test with the bbox constants as f64 literals                    (:62-78)
 *   arithmetic near/far swap by |t|                                      (:81-83)
 *   start/end = centre + t*dir in f32                                    (:86-89)
 * bbox is passed as f32[6] and promoted to double (the reference substitutes the
 * decimal text of np.float32 values into the source, raynet_fp.py:241-246).
 * ---------------------------------------------------------------------------------- */
RN_EXPORT void rn_oracle_sample_in_bbox(int ray_idx, int H, const float *P_inv, const float *centre,
                                        const float *bbox, float *ray_start, float *ray_end) {
    float pixel[2];
    pixel[0] = (float)(ray_idx / H);
    pixel[1] = (float)(ray_idx % H);

    double out[3], normalizer;
    normalizer = out[0] = out[1] = out[2] = 0;
    for (int r = 0; r < 3; r++) {
        out[r] += (double)(P_inv[r * 3 + 0] * pixel[0]);
        out[r] += (double)(P_inv[r * 3 + 1] * pixel[1]);
        out[r] += (double)P_inv[r * 3 + 2] * 1.0;
    }
    normalizer += (double)(P_inv[3 * 3 + 0] * pixel[0]);
    normalizer += (double)(P_inv[3 * 3 + 1] * pixel[1]);
    normalizer += (double)P_inv[3 * 3 + 2] * 1.0;
    out[0] /= normalizer;
    out[1] /= normalizer;
    out[2] /= normalizer;

    float dir[3];
    for (int i = 0; i < 3; i++) dir[i] = (float)(out[i] - (double)centre[i]);

    float t_near = -INFINITY, t_far = INFINITY, t1, t2;
    for (int a = 0; a < 3; a++) {
        t1 = (float)(((double)bbox[a] - (double)centre[a]) / (double)dir[a]);
        t2 = (float)(((double)bbox[3 + a] - (double)centre[a]) / (double)dir[a]);
        t_near = fmaxf(fminf(t1, t2), t_near);
        t_far = fminf(fmaxf(t1, t2), t_far);
    }
    float near_mask = (fabsf(t_near) < fabsf(t_far)) ? 1.0f : 0.0f;
    float t_near_actual = t_near * near_mask + t_far * (1 - near_mask);
    float t_far_actual = (1 - near_mask) * t_near + near_mask * t_far;
    for (int i = 0; i < 3; i++) {
        ray_start[i] = centre[i] + t_near_actual * dir[i];
        ray_end[i] = centre[i] + t_far_actual * dir[i];
    }
}

/* ------------------------------------------------------------------------------------
 * a2. similarity -- cuda_implementations/feature_similarities.cu:10-124
 *   plane point k: start + k*(end-start)/(D-1) in f32                    (:84-86)
 *   projection with 3x4 P, all f32, sequential adds                      (:10-32)
 *   pixel -> feature index: roundf + padding - (padding-1)/2, clamp x to [0,W],
 *   y to [0,H]; if either is 0 both become 0                             (:42-61)
 *   feature address view*(H+p+1)(W+p+1)F + y*(W+p+1)F + x*F              (:74-78,94-96)
 *   S[k] += dot(f_i, f_j), sequential f32 sum over F                     (:34-40,99)
 *   S[k] /= (V*(V-1))/2 (integer)                                        (:105-107)
 *   stable softmax over D                                                (:109-123)
 * S is zero-initialised here (reference defect: the fused kernels leave it
 * uninitialised, raynet_fp.py:77; the non-fused caller zero-fills, forward_pass.py:320).
 * ---------------------------------------------------------------------------------- */
static inline void dot_m34v3(const float *m, const float *v, float *out) {
    float normalizer;
    normalizer = out[0] = out[1] = 0;
    out[0] += m[0 * 4 + 0] * v[0];
    out[0] += m[0 * 4 + 1] * v[1];
    out[0] += m[0 * 4 + 2] * v[2];
    out[0] += m[0 * 4 + 3] * 1;
    out[1] += m[1 * 4 + 0] * v[0];
    out[1] += m[1 * 4 + 1] * v[1];
    out[1] += m[1 * 4 + 2] * v[2];
    out[1] += m[1 * 4 + 3] * 1;
    normalizer += m[2 * 4 + 0] * v[0];
    normalizer += m[2 * 4 + 1] * v[1];
    normalizer += m[2 * 4 + 2] * v[2];
    normalizer += m[2 * 4 + 3] * 1;
    out[0] /= normalizer;
    out[1] /= normalizer;
}

static inline void pixel_to_features(const float *x, int *f_idx, int padding, int h, int w) {
    f_idx[0] = (int)(roundf(x[0]) + (float)(padding - (padding - 1) / 2));
    f_idx[1] = (int)(roundf(x[1]) + (float)(padding - (padding - 1) / 2));
    f_idx[0] = f_idx[0] > 0 ? f_idx[0] : 0;
    f_idx[0] = f_idx[0] < w ? f_idx[0] : w;
    f_idx[1] = f_idx[1] > 0 ? f_idx[1] : 0;
    f_idx[1] = f_idx[1] < h ? f_idx[1] : h;
    if (f_idx[0] == 0 || f_idx[1] == 0) f_idx[0] = f_idx[1] = 0;
}

RN_EXPORT void rn_oracle_similarity(const float *features, const float *P, const float *ray_start,
                                    const float *ray_end, int D, int V, int F, int H, int W, int padding,
                                    float *S) {
    const int64_t fh = H + padding + 1, fw = W + padding + 1;
    const int64_t dim_x = fh * fw * F, dim_y = fw * F, dim_z = F;
    float pixel_i[2], pixel_j[2], point[3];
    int f_idx[2];
    for (int k = 0; k < D; k++) S[k] = 0.0f;
    for (int i = 0; i < V; i++) {
        for (int j = i + 1; j < V; j++) {
            for (int k = 0; k < D; k++) {
                for (int a = 0; a < 3; a++)
                    point[a] = ray_start[a] + (float)k * (ray_end[a] - ray_start[a]) / (float)(D - 1);
                dot_m34v3(P + i * 12, point, pixel_i);
                dot_m34v3(P + j * 12, point, pixel_j);
                pixel_to_features(pixel_i, f_idx, padding, H, W);
                const float *fi = features + dim_x * i + dim_y * f_idx[1] + dim_z * f_idx[0];
                pixel_to_features(pixel_j, f_idx, padding, H, W);
                const float *fj = features + dim_x * j + dim_y * f_idx[1] + dim_z * f_idx[0];
                float sum = 0.0f;
                for (int c = 0; c < F; c++) sum += fi[c] * fj[c];
                S[k] += sum;
            }
        }
    }
    const int npairs = (V * (V - 1)) / 2;
    for (int k = 0; k < D; k++) S[k] /= (float)npairs;
    float maximum = -INFINITY;
    for (int k = 0; k < D; k++) maximum = fmaxf(maximum, S[k]);
    float sum = 0.0f;
    for (int k = 0; k < D; k++) {
        S[k] = expf(S[k] - maximum);
        sum += S[k];
    }
    for (int k = 0; k < D; k++) S[k] /= sum;
}

/* Feature-map pixel (x, y) that plane point k of a ray lands on in one view: the integer
 * decision inside a2, exposed so tests can gate it bit-exactly. */
RN_EXPORT void rn_oracle_project_pixel(const float *P_view, const float *ray_start, const float *ray_end,
                                       int k, int D, int H, int W, int padding, int *f_idx) {
    float point[3], px[2];
    for (int a = 0; a < 3; a++)
        point[a] = ray_start[a] + (float)k * (ray_end[a] - ray_start[a]) / (float)(D - 1);
    dot_m34v3(P_view, point, px);
    pixel_to_features(px, f_idx, padding, H, W);
}

/* ------------------------------------------------------------------------------------
 * a3. voxel_traversal -- ray_marching/ray_tracing.pyx:64-199 (the Cython flavour:
 * run-time f32 bbox and int32 grid shape, bin = (max-min)/grid in f32), identical in
 * structure to cuda_implementations/ray_tracing.cu:9-143.
 * Returns the number of traversed voxels; writes int32 triplets to voxels[N][3].
 * ---------------------------------------------------------------------------------- */
RN_EXPORT int rn_oracle_voxel_traversal(const float *bbox, const int *grid, int *voxels, int N,
                                        const float *ray_start, const float *ray_end) {
    const float EPS = 1e-2f;
    float s[3], e[3], bin[3], ray[3], tMax[3], tDelta[3];
    int step[3], cur[3], last[3];
    for (int a = 0; a < 3; a++) {
        s[a] = ray_start[a] - bbox[a];            /* pyx:101 */
        e[a] = ray_end[a] - bbox[a];              /* pyx:102 */
        bin[a] = bbox[3 + a] - bbox[a];           /* pyx:103 */
        bin[a] = bin[a] / (float)grid[a];         /* pyx:104 */
    }
    for (int a = 0; a < 3; a++) {
        ray[a] = e[a] - s[a];                     /* pyx:107 */
        step[a] = ray[a] >= 0 ? 1 : -1;           /* pyx:108-110 */
    }
    for (int a = 0; a < 3; a++) {
        /* pyx:114-119: stepX*bin_size[0]*_EPS evaluated left to right in f32 */
        float nudge = ((float)step[a] * bin[a]) * EPS;
        s[a] = s[a] + nudge;
        e[a] = e[a] - nudge;
    }
    for (int a = 0; a < 3; a++) {
        cur[a] = (int)floor((double)(s[a] / bin[a]));   /* pyx:35-39, libc floor on the f32 quotient */
        last[a] = (int)floor((double)(e[a] / bin[a]));
    }
    if (!(cur[0] >= 0 && cur[0] < grid[0] && cur[1] >= 0 && cur[1] < grid[1] && cur[2] >= 0 &&
          cur[2] < grid[2]))
        return 0;                                  /* pyx:125-126 */
    for (int a = 0; a < 3; a++) {
        tMax[a] = FLT_MAX;                         /* pyx:130 */
        if (ray[a] != 0) {
            float cc = (float)cur[a] * bin[a];     /* pyx:133 */
            float b;
            if (step[a] < 0 && cc < s[a])
                b = cc;
            else
                b = cc + (float)step[a] * bin[a];
            tMax[a] = (b - s[a]) / ray[a];         /* pyx:139 */
        }
        tDelta[a] = (ray[a] != 0) ? ((float)step[a] * bin[a]) / ray[a] : FLT_MAX; /* pyx:161-163 */
    }
    int ii = 0;
    voxels[0] = cur[0]; voxels[1] = cur[1]; voxels[2] = cur[2];
    ii = 1;
    while (!(cur[0] == last[0] && cur[1] == last[1] && cur[2] == last[2]) && ii < N) {
        int a;
        if (tMax[0] < tMax[1])
            a = (tMax[0] < tMax[2]) ? 0 : 2;       /* pyx:169-181 */
        else
            a = (tMax[1] < tMax[2]) ? 1 : 2;       /* pyx:183-194 */
        cur[a] += step[a];
        if (cur[a] < 0 || cur[a] >= grid[a]) return ii;
        tMax[a] = tMax[a] + tDelta[a];
        voxels[3 * ii + 0] = cur[0];
        voxels[3 * ii + 1] = cur[1];
        voxels[3 * ii + 2] = cur[2];
        ii++;
    }
    return ii;
}

/* Voxel centres exactly as utils/generic_utils.py:104-110 (get_voxel_grid):
 *   x_i = f32( f64(min) + i * (f64(max)-f64(min))/G )   (np.linspace works in f64)
 *   bin = x_1 - x_0 in f32;  centre = x_i + bin/2 in f32.
 * Layout out[Gx][Gy][Gz][3], i.e. the transpose(1,2,3,0) of forward_pass.py:571-576. */
RN_EXPORT void rn_oracle_voxel_grid(const float *bbox, const int *grid, float *out) {
    float *ax[3];
    float half[3];
    for (int a = 0; a < 3; a++) {
        ax[a] = (float *)malloc(sizeof(float) * (size_t)grid[a]);
        double start = (double)bbox[a], stop = (double)bbox[3 + a];
        double step = (stop - start) / (double)grid[a];
        for (int i = 0; i < grid[a]; i++) ax[a][i] = (float)((double)i * step + start);
        float b = (grid[a] > 1) ? (ax[a][1] - ax[a][0]) : (float)step;
        half[a] = b / 2;
    }
    for (int x = 0; x < grid[0]; x++)
        for (int y = 0; y < grid[1]; y++)
            for (int z = 0; z < grid[2]; z++) {
                float *o = out + (((size_t)x * grid[1] + y) * grid[2] + z) * 3;
                o[0] = ax[0][x] + half[0];
                o[1] = ax[1][y] + half[1];
                o[2] = ax[2][z] + half[2];
            }
    for (int a = 0; a < 3; a++) free(ax[a]);
}

/* ------------------------------------------------------------------------------------
 * a4. planes_voxels_mapping -- cuda_implementations/planes_voxels_mapping.cu:6-92
 * (numerically the same interpolation as planes_voxels_mapping.py:165-211 li_2).
 * ---------------------------------------------------------------------------------- */
RN_EXPORT void rn_oracle_planes_voxels_mapping(const float *voxel_grid, const int *grid,
                                               const int *ray_voxel_indices, int count,
                                               const float *ray_start, const float *ray_end,
                                               const float *S, int D, float *S_new) {
    float sum = 0.0f, eps = 1e-4f;
    float ray[3];
    for (int i = 0; i < 3; i++) ray[i] = ray_end[i] - ray_start[i];
    float ray_norm = 0.0f;
    for (int i = 0; i < 3; i++) ray_norm += ray[i] * ray[i];
    float vd, t, left_d, right_d, coeff_1, coeff_2;
    float start = 0.0f, end = 1.0f;
    float step = (end - start) / (float)(D - 1);
    int left = 0, right = 1;
    const int64_t dim_x = 3 * (int64_t)grid[1] * grid[2], dim_y = 3 * grid[2], dim_z = 3;
    float srsum = 0.0f;
    for (int i = 0; i < count; i++) {
        sum = 0.0f;
        int ix = ray_voxel_indices[3 * i], iy = ray_voxel_indices[3 * i + 1], iz = ray_voxel_indices[3 * i + 2];
        for (int j = 0; j < 3; j++) {
            vd = voxel_grid[ix * dim_x + iy * dim_y + iz * dim_z + j];
            vd -= ray_start[j];
            sum += ray[j] * vd;
        }
        t = clampf(sum / ray_norm, eps, 1 - eps);
        left_d = t - (start + (float)left * step);
        right_d = t - (start + (float)right * step);
        while (left_d > 0 && right_d > 0) {
            left++;
            right++;
            left_d = t - (start + (float)left * step);
            right_d = t - (start + (float)right * step);
        }
        left_d = fabsf(left_d);
        right_d = fabsf(right_d);
        coeff_1 = (float)(1.0 - (double)(left_d / (left_d + right_d)));
        coeff_2 = (float)(1.0 - (double)(right_d / (left_d + right_d)));
        S_new[i] = coeff_1 * S[left] + coeff_2 * S[right];
        srsum += S_new[i];
    }
    for (int i = 0; i < count; i++) S_new[i] = S_new[i] / srsum;
}

/* ------------------------------------------------------------------------------------
 * a5-a8. ray-potential BP -- mrf/mrf_np.py (the runnable reference backend).
 *
 * Precision follows NumPy's promotion rules statement by statement.  One switch:
 * acc_f64 = 1 reproduces the reference *as it executes under NumPy >= 2* (the
 * accumulators np.ones(f32) * np.float64 become float64, mrf_np.py:285-292, so the
 * occupancy-to-ray chain is f64); acc_f64 = 0 reproduces NumPy < 2 (2018) and the
 * reference's CUDA/TF backends, where the accumulators and that chain are f32.
 * ---------------------------------------------------------------------------------- */
static inline int64_t grid_lin(const int *grid, const int *v) {
    return ((int64_t)v[0] * grid[1] + v[1]) * grid[2] + v[2];
}

/* mrf_np.py:4-8 : np.clip(x, eps, 1-eps) on f32, x / x.sum() in f32 */
static void clip_and_renorm(const float *S, int c, float *s) {
    const float lo = (float)1e-5, hi = (float)(1 - 1e-5);
    double sum = 0.0; /* NumPy pairwise f32 sum is accurate to ~1 ulp; f64 accumulate then round */
    for (int i = 0; i < c; i++) {
        s[i] = fminf(fmaxf(S[i], lo), hi);
        sum += (double)s[i];
    }
    float fs = (float)sum;
    for (int i = 0; i < c; i++) s[i] = s[i] / fs;
}

/* mrf_np.py:50-80 : o_i (clipped sigmoid) and exclusive cumprod of (1-o).  Results are
 * returned as doubles holding values of the working precision. */
static void occupancy_to_ray(const void *acc, int acc_f64, const int *grid, const int *vox, const float *msg,
                             int c, double *o, double *cp /* c+1 */) {
    cp[0] = 1.0;
    if (acc_f64) {
        const double *A = (const double *)acc;
        double run = 1.0;
        for (int i = 0; i < c; i++) {
            double x = A[grid_lin(grid, vox + 3 * i)] - (double)msg[i];
            double m = fmax(0.0, x);
            double t1 = exp(0.0 - m), t2 = exp(x - m);
            double v = t2 / (t2 + t1);
            v = fmin(fmax(v, 1e-4), 1 - 1e-4);
            o[i] = v;
            run = run * (1 - v);
            cp[i + 1] = run;
        }
    } else {
        const float *A = (const float *)acc;
        float run = 1.0f;
        const float lo = (float)1e-4, hi = (float)(1 - 1e-4);
        for (int i = 0; i < c; i++) {
            float x = A[grid_lin(grid, vox + 3 * i)] - msg[i];
            float m = fmaxf(0.0f, x);
            float t1 = expf(0.0f - m), t2 = expf(x - m);
            float v = t2 / (t2 + t1);
            v = fminf(fmaxf(v, lo), hi);
            o[i] = (double)v;
            run = run * (1 - v);   /* f32 cumprod */
            cp[i + 1] = (double)run;
        }
    }
}

/* mrf_np.py:11-126 single_ray_belief_propagation.  s = clip_and_renorm(S[r,:c]).
 * scratch: 4*(c+1) doubles.  Output t[c] (f32 messages). */
static void single_ray_bp(const void *acc, int acc_f64, const int *grid, const int *vox, const float *msg,
                          const float *s, int c, double *scratch, float *t) {
    double *o = scratch, *cp = scratch + (c + 1), *a = scratch + 2 * (c + 1), *suf = scratch + 3 * (c + 1);
    occupancy_to_ray(acc, acc_f64, grid, vox, msg, c, o, cp);
    for (int i = 0; i < c; i++) a[i] = (o[i] * cp[i]) * (double)s[i];   /* :91, :110 (f64) */
    /* :109-112 reversed cumsum of hstack([a, 0]) reversed again, [1:] */
    double run = 0.0;
    for (int i = c - 1; i >= 0; i--) {
        suf[i] = run;           /* sum_{j>i} a_j, accumulated from the far end like the reversed cumsum */
        run = run + a[i];
    }
    double pre = 0.0;
    for (int i = 0; i < c; i++) {
        /* :90-92 both rows get f32(0 + cumsum) */
        float pre32 = (float)pre;
        /* :95 row1 += cp*s ; :109-112 row0 += suffix/(1-o)  (f64 add, stored f32) */
        float pos = (float)((double)pre32 + cp[i] * (double)s[i]);
        double one_minus_o = acc_f64 ? (1 - o[i]) : (double)(1 - (float)o[i]);
        float neg = (float)((double)pre32 + suf[i] / one_minus_o);
        float p = pos / (pos + neg);                         /* :115-116 f32 */
        t[i] = logf(p) - logf(1 - p);                        /* :120 */
        pre = pre + a[i];
    }
}

/* mrf_np.py:129-203 single_ray_depth_estimate -> P / P.sum() (f64), stored f32 by the caller (:378). */
static void single_ray_depth(const void *acc, int acc_f64, const int *grid, const int *vox, const float *msg,
                             const float *s, int c, double *scratch, float *out) {
    double *o = scratch, *cp = scratch + (c + 1), *a = scratch + 2 * (c + 1);
    occupancy_to_ray(acc, acc_f64, grid, vox, msg, c, o, cp);
    double sum = 0.0;
    for (int i = 0; i < c; i++) {
        a[i] = (o[i] * cp[i]) * (double)s[i];
        sum += a[i];
    }
    for (int i = 0; i < c; i++) out[i] = (float)(a[i] / sum);
}

/* One synchronous BP sweep over all rays (the body of the `for it` loop, mrf_np.py:297-315):
 * reads acc_prev + msgs, writes msgs in place, adds into acc_new.  Rays with count <= 1 are
 * skipped (:299-301).  acc arrays are f64 if acc_f64 else f32.  OpenMP over rays; the
 * scatter-add uses atomics (summation order is not defined in the reference either). */
RN_EXPORT void rn_oracle_bp_iteration(const float *S, const int *ray_voxel_indices, const int *ray_voxel_count,
                                      int64_t N, int M, const int *grid, const void *acc_prev, void *acc_new,
                                      int acc_f64, float *msgs) {
#pragma omp parallel
    {
        double *scratch = (double *)malloc(sizeof(double) * 4 * (size_t)(M + 1));
        float *s = (float *)malloc(sizeof(float) * (size_t)M);
        float *t = (float *)malloc(sizeof(float) * (size_t)M);
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < N; r++) {
            int c = ray_voxel_count[r];
            if (c <= 1) continue;
            const int *vox = ray_voxel_indices + r * (int64_t)M * 3;
            clip_and_renorm(S + r * (int64_t)M, c, s);
            single_ray_bp(acc_prev, acc_f64, grid, vox, msgs + r * (int64_t)M, s, c, scratch, t);
            for (int i = 0; i < c; i++) {
                int64_t g = grid_lin(grid, vox + 3 * i);
                if (acc_f64) {
#pragma omp atomic
                    ((double *)acc_new)[g] += (double)t[i];
                } else {
#pragma omp atomic
                    ((float *)acc_new)[g] += t[i];
                }
                msgs[r * (int64_t)M + i] = t[i];
            }
        }
        free(scratch); free(s); free(t);
    }
}

/* mrf_np.py:243-330 belief_propagation: msgs = 0; both accumulators = log(g) - log(1-g);
 * I sweeps with swap + refill.  acc_out receives ray_to_occupancy_accumulated_prev_pon. */
RN_EXPORT void rn_oracle_belief_propagation(const float *S, const int *ray_voxel_indices,
                                            const int *ray_voxel_count, int64_t N, int M, const int *grid,
                                            double gamma, int bp_iterations, int acc_f64, float *msgs,
                                            void *acc_out) {
    size_t G = (size_t)grid[0] * grid[1] * grid[2];
    size_t esz = acc_f64 ? sizeof(double) : sizeof(float);
    double prior = log(gamma) - log(1 - gamma);
    void *prev = acc_out, *nw = malloc(G * esz);
    memset(msgs, 0, sizeof(float) * (size_t)N * M);
    for (size_t i = 0; i < G; i++) {
        if (acc_f64) { ((double *)prev)[i] = prior; ((double *)nw)[i] = prior; }
        else { ((float *)prev)[i] = (float)prior; ((float *)nw)[i] = (float)prior; }
    }
    for (int it = 0; it < bp_iterations; it++) {
        rn_oracle_bp_iteration(S, ray_voxel_indices, ray_voxel_count, N, M, grid, prev, nw, acc_f64, msgs);
        memcpy(prev, nw, G * esz);                                    /* :318 */
        for (size_t i = 0; i < G; i++) {                              /* :319 */
            if (acc_f64) ((double *)nw)[i] = prior; else ((float *)nw)[i] = (float)prior;
        }
    }
    free(nw);
}

/* mrf_np.py:333-385 compute_depth_distribution: S_new zero-filled, rays with count<=1 skipped. */
RN_EXPORT void rn_oracle_depth_distribution(const float *S, const int *ray_voxel_indices,
                                            const int *ray_voxel_count, int64_t N, int M, const int *grid,
                                            const void *acc, int acc_f64, const float *msgs, float *S_new) {
    memset(S_new, 0, sizeof(float) * (size_t)N * M);
#pragma omp parallel
    {
        double *scratch = (double *)malloc(sizeof(double) * 4 * (size_t)(M + 1));
        float *s = (float *)malloc(sizeof(float) * (size_t)M);
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < N; r++) {
            int c = ray_voxel_count[r];
            if (c <= 1) continue;
            clip_and_renorm(S + r * (int64_t)M, c, s);
            single_ray_depth(acc, acc_f64, grid, ray_voxel_indices + r * (int64_t)M * 3, msgs + r * (int64_t)M,
                             s, c, scratch, S_new + r * (int64_t)M);
        }
        free(scratch); free(s);
    }
}

/* a10. mrf_np.py:206-240 compute_occupancy_probabilities: max-shifted sigmoid, in acc's dtype. */
RN_EXPORT void rn_oracle_occupancy(const void *acc, int acc_f64, int64_t G, float *out) {
    for (int64_t i = 0; i < G; i++) {
        if (acc_f64) {
            double x = ((const double *)acc)[i];
            double m = fmax(0.0, x), t1 = exp(0.0 - m), t2 = exp(x - m);
            out[i] = (float)(t2 / (t2 + t1));
        } else {
            float x = ((const float *)acc)[i];
            float m = fmaxf(0.0f, x), t1 = expf(0.0f - m), t2 = expf(x - m);
            out[i] = t2 / (t2 + t1);
        }
    }
}

/* a9. arg-max -> depth -- cuda_implementations/raynet_fp.py:193-226.  First maximum over
 * ALL M slots (strict > from -inf), voxel centre of that slot, distance to the camera. */
RN_EXPORT float rn_oracle_argmax_depth(const float *S_row, int M, const int *ray_voxel_indices,
                                       const float *voxel_grid, const int *grid, const float *centre,
                                       int *argmax_out) {
    float mx = -INFINITY;
    int max_idx = 0;
    for (int i = 0; i < M; i++)
        if (S_row[i] > mx) { max_idx = i; mx = S_row[i]; }
    const int *v = ray_voxel_indices + 3 * max_idx;
    const float *p = voxel_grid + grid_lin(grid, v) * 3;
    float sum = 0.0f;
    for (int i = 0; i < 3; i++) {
        float d = p[i] - centre[i];
        sum += d * d;
    }
    if (argmax_out) *argmax_out = max_idx;
    return sqrtf(sum);
}

/* ------------------------------------------------------------------------------------
 * Batched drivers (OpenMP over rays) used by tests and by the CPU baseline.
 * ---------------------------------------------------------------------------------- */

/* Front end for a batch of rays of ONE reference image, = the device function
 * mvcnn_ray_marching_with_voxels_mapping of raynet_fp.py:55-104:
 * sample_in_bbox -> similarity -> DDA -> plane->voxel.  Outputs are the dense reference
 * buffers: idx[N][M][3] (zero beyond count), cnt[N], S_vox[N][M] (zero beyond count),
 * plus start/end[N][3] and S[N][D] for stage-wise gates (may be NULL). */
RN_EXPORT void rn_oracle_frontend(const int *ray_idxs, int64_t N, const float *features, const float *P,
                                  const float *P_inv, const float *centre, const float *voxel_grid,
                                  const float *bbox, const int *grid, int M, int D, int V, int F, int H, int W,
                                  int padding, int *idx, int *cnt, float *S_vox, float *starts, float *ends,
                                  float *S_planes) {
#pragma omp parallel
    {
        float *S = (float *)malloc(sizeof(float) * (size_t)D);
#pragma omp for schedule(dynamic, 16)
        for (int64_t r = 0; r < N; r++) {
            float rs[3], re[3];
            rn_oracle_sample_in_bbox(ray_idxs[r], H, P_inv, centre, bbox, rs, re);
            rn_oracle_similarity(features, P, rs, re, D, V, F, H, W, padding, S);
            int *vi = idx + r * (int64_t)M * 3;
            float *sv = S_vox + r * (int64_t)M;
            memset(vi, 0, sizeof(int) * (size_t)M * 3);
            memset(sv, 0, sizeof(float) * (size_t)M);
            int c = rn_oracle_voxel_traversal(bbox, grid, vi, M, rs, re);
            cnt[r] = c;
            if (c > 0) rn_oracle_planes_voxels_mapping(voxel_grid, grid, vi, c, rs, re, S, D, sv);
            if (starts) memcpy(starts + 3 * r, rs, sizeof rs);
            if (ends) memcpy(ends + 3 * r, re, sizeof re);
            if (S_planes) memcpy(S_planes + r * (int64_t)D, S, sizeof(float) * (size_t)D);
        }
        free(S);
    }
}

RN_EXPORT void rn_oracle_batch_sample_in_bbox(const int *ray_idxs, int64_t N, int H, const float *P_inv,
                                              const float *centre, const float *bbox, float *starts,
                                              float *ends) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; r++)
        rn_oracle_sample_in_bbox(ray_idxs[r], H, P_inv, centre, bbox, starts + 3 * r, ends + 3 * r);
}

RN_EXPORT void rn_oracle_batch_similarity(const float *features, const float *P, const float *starts,
                                          const float *ends, int64_t N, int D, int V, int F, int H, int W,
                                          int padding, float *S) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t r = 0; r < N; r++)
        rn_oracle_similarity(features, P, starts + 3 * r, ends + 3 * r, D, V, F, H, W, padding, S + r * D);
}

RN_EXPORT void rn_oracle_batch_voxel_traversal(const float *bbox, const int *grid, const float *starts,
                                               const float *ends, int64_t N, int M, int *idx, int *cnt) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < N; r++)
        cnt[r] = rn_oracle_voxel_traversal(bbox, grid, idx + r * (int64_t)M * 3, M, starts + 3 * r, ends + 3 * r);
}

RN_EXPORT void rn_oracle_batch_planes_voxels_mapping(const float *voxel_grid, const int *grid, const int *idx,
                                                     const int *cnt, const float *starts, const float *ends,
                                                     const float *S, int64_t N, int M, int D, float *S_new) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < N; r++)
        if (cnt[r] > 0)
            rn_oracle_planes_voxels_mapping(voxel_grid, grid, idx + r * (int64_t)M * 3, cnt[r], starts + 3 * r,
                                            ends + 3 * r, S + r * (int64_t)D, D, S_new + r * (int64_t)M);
}

RN_EXPORT void rn_oracle_batch_argmax_depth(const float *S_new, const int *idx, int64_t N, int M,
                                            const float *voxel_grid, const int *grid, const float *centre,
                                            float *depth, int *argmax) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; r++)
        depth[r] = rn_oracle_argmax_depth(S_new + r * (int64_t)M, M, idx + r * (int64_t)M * 3, voxel_grid, grid,
                                          centre, argmax ? argmax + r : NULL);
}
