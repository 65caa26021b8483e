/*
 * raynet_b200.h -- C-ABI of the B200-native RayNet volumetric-inference hot path.
 *
 * Drop-in boundary: every entry point replaces one PyCUDA `prepared_call` site of the
 * reference (paths relative to /root/reference/raynet).  Conventions:
 *   - all pointers are DEVICE pointers (what PyCUDA passes as `.gpudata`), C-contiguous;
 *   - the caller allocates every output;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); launches are
 *     asynchronous on it and correct under the reference's fully synchronous usage;
 *   - sizes that the reference text-substitutes into the CUDA source at JIT time
 *     (raynet_fp.py:230-248) travel in an RnParams POD at launch time -- no JIT;
 *   - return value: 0 on success, otherwise an RnStatus / cudaError_t code; the message is
 *     available from rn_last_error().  The Python shim raises AssertionError for
 *     RN_ERR_SHAPE (the reference asserts, raynet_fp.py:290-301) and RuntimeError else.
 *
 * Layouts are the reference's: ray ids are column-major pixel indices id = x*H + y
 * (sampling_schemes.cu:5-8); ray_voxel_indices is int32 [B][M][3]; S / messages are
 * float32 [B][M]; accumulators are float32 [Gx][Gy][Gz]; features are float32
 * [V][H+p+1][W+p+1][F]; P is float32 [V][3][4]; P_inv float32 [4][3]; centre float32 [4].
 */
#ifndef RAYNET_B200_H
#define RAYNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RN_ABI_VERSION 3

typedef enum RnStatus {
    RN_OK = 0,
    RN_ERR_SHAPE = 10001,     /* a size/shape the reference would `assert` on */
    RN_ERR_UNSUPPORTED = 10002, /* e.g. max_voxels above the compiled limit */
    RN_ERR_CUDA = 10003       /* a CUDA runtime error; see rn_last_error() */
} RnStatus;

/* Compile-time constants of the reference's templated kernels, as run-time data.
 * Mirrors the arguments of perform_raynet_fp (cuda_implementations/raynet_fp.py:10-21). */
typedef struct RnParams {
    int32_t max_voxels;    /* M: maximum number of marched voxels per ray            */
    int32_t depth_planes;  /* D: discretisation steps along the ray                  */
    int32_t n_views;       /* N: number of views = neighbors + 1 (reference first)   */
    int32_t feat_dim;      /* F: feature size of the multi-view CNN                  */
    int32_t height;        /* H: image height                                        */
    int32_t width;         /* W: image width                                         */
    int32_t padding;       /* zero-padding of the CNN input (generation_parameters)  */
    int32_t grid[3];       /* voxel grid shape (Gx, Gy, Gz)                          */
    float bbox[6];         /* (min_x, min_y, min_z, max_x, max_y, max_z), float32    */
} RnParams;

const char *rn_last_error(void);
int rn_abi_version(void);
/* Number of SMs / name of the current device (diagnostics for bench.py). */
int rn_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ---------------------------------------------------------------------------------------
 * Stage-wise entry points (one per reference kernel)
 * ------------------------------------------------------------------------------------- */

/* sample_in_bbox for a batch of rays: ray start / end on the bbox.
 * Replaces the device function sampling_schemes.cu:44-90 (no stand-alone kernel in the
 * reference; exposed because every fused kernel starts with it).
 * out: starts, ends float32 [n][3]. */
int rn_sample_in_bbox(const RnParams *p, const int32_t *ray_idxs, const float *P_inv, const float *centre,
                      float *starts, float *ends, int64_t n_rays, void *stream);

/* batch_sample_points_in_bbox (sampling_schemes.cu:92-122; wrapped by
 * cuda_implementations/sample_points.py:12-54): D uniformly spaced homogeneous points
 * per ray.  out: points float32 [n][D][4]. */
int rn_sample_points(const RnParams *p, const int32_t *ray_idxs, const float *P_inv, const float *centre,
                     float *points, int64_t n_rays, void *stream);

/* batch_compute_similarities (feature_similarities.cu:126-146): per-ray softmax depth
 * distribution over D planes from V feature maps.  out: S float32 [n][D] (overwritten;
 * the reference accumulates into a pre-zeroed S, forward_pass.py:320). */
int rn_similarity(const RnParams *p, const float *features, const float *P, const float *starts,
                  const float *ends, float *S, int64_t n_rays, void *stream);

/* batch_multi_view_cnn_forward_pass (similarities.py:46-79): sample_in_bbox + similarity. */
int rn_mvcnn_forward(const RnParams *p, const int32_t *ray_idxs, const float *features, const float *P,
                     const float *P_inv, const float *centre, float *S, int64_t n_rays, void *stream);

/* batch_multi_view_cnn_forward_pass_with_depth (similarities.py:168-230): additionally the
 * D points per ray and depth_map[r] = |point[argmax S] - centre|.
 * out: S [n][D], points [n][D][4], depth_map [n]. */
int rn_mvcnn_forward_depth(const RnParams *p, const int32_t *ray_idxs, const float *features, const float *P,
                           const float *P_inv, const float *centre, float *S, float *points,
                           float *depth_map, int64_t n_rays, void *stream);

/* batch_voxel_traversal (ray_tracing.cu:145-163; wrapped by ray_marching/ray_tracing_cuda.py:12-63).
 * Arithmetic follows the Cython flavour ray_tracing.pyx:99-199 (run-time float32 bbox).
 * out: ray_voxel_indices int32 [n][M][3] (only the first count triplets are written),
 *      ray_voxel_count int32 [n] (always written, 0 for rays starting outside). */
int rn_voxel_traversal(const RnParams *p, const float *starts, const float *ends, int32_t *ray_voxel_indices,
                       int32_t *ray_voxel_count, int64_t n_rays, void *stream);

/* batch_planes_voxels_mapping (planes_voxels_mapping.cu:94-118; wrapped by
 * planes_voxels_mapping/planes_voxels_mapping_cuda.py:11-67).
 * voxel_grid float32 [Gx][Gy][Gz][3] exactly as the reference passes it.
 * out: S_new float32 [n][M] (first count entries written). */
int rn_planes_to_voxels(const RnParams *p, const float *voxel_grid, const int32_t *ray_voxel_indices,
                        const int32_t *ray_voxel_count, const float *starts, const float *ends, const float *S,
                        float *S_new, int64_t n_rays, void *stream);

/* batch_belief_propagation (mrf_bp.cu:180-204; wrapped by mrf/mrf_cuda.py:37-79): one
 * synchronous sweep.  Reads acc_in + msgs, writes msgs IN PLACE (every reference caller
 * aliases in/out, mrf_cuda.py:73-75) and atomically adds the new messages into acc_out.
 * Follows mrf_np.py: S is clipped/renormalised on the fly and NOT modified; rays with
 * count <= 1 are skipped (mrf_np.py:299-301). */
int rn_bp_iteration(const RnParams *p, const float *S, const int32_t *ray_voxel_indices,
                    const int32_t *ray_voxel_count, const float *acc_in, float *msgs, float *acc_out,
                    int64_t n_rays, void *stream);

/* batch_depth_estimation (mrf_bp.cu:206-228; mrf_cuda.py:81-122): S_new float32 [n][M];
 * entries >= count and rays with count <= 1 are set to 0 (mrf_np.py:370-383). */
int rn_depth_estimate(const RnParams *p, const float *S, const int32_t *ray_voxel_indices,
                      const int32_t *ray_voxel_count, const float *acc, const float *msgs, float *S_new,
                      int64_t n_rays, void *stream);

/* compute_occupancy_probabilities (mrf_np.py:206-240): out[i] = sigmoid(acc[i]). */
int rn_occupancy(const float *acc, float *out, int64_t n, void *stream);

/* GPUArray.fill replacement used between sweeps (forward_pass.py:676-678). */
int rn_fill_f32(float *dst, float value, int64_t n, void *stream);

/* ---------------------------------------------------------------------------------------
 * Fused entry points (the reference's hot kernels)
 * ------------------------------------------------------------------------------------- */

/* batch_raynet_fp (raynet_fp.py:106-149): sample -> similarity -> DDA -> plane->voxel ->
 * one BP sweep.  Fills ray_voxel_indices, ray_voxel_count, S_voxel_space (the reference's
 * scratch outputs), updates msgs in place and adds into acc_out. */
int rn_raynet_fp(const RnParams *p, const int32_t *ray_idxs, const float *features, const float *P,
                 const float *P_inv, const float *centre, const float *voxel_grid,
                 int32_t *ray_voxel_indices, int32_t *ray_voxel_count, float *S_voxel_space,
                 const float *acc_in, float *msgs, float *acc_out, int64_t n_rays, void *stream);

/* batch_complete_depth_estimation (raynet_fp.py:151-227): same front end -> depth
 * re-estimation (written over S_voxel_space) -> arg-max voxel -> distance to the camera.
 * out: depth_map float32 [n]. */
int rn_raynet_de(const RnParams *p, const int32_t *ray_idxs, const float *features, const float *P,
                 const float *P_inv, const float *centre, const float *voxel_grid,
                 int32_t *ray_voxel_indices, int32_t *ray_voxel_count, float *S_voxel_space,
                 const float *acc, const float *msgs, float *depth_map, int64_t n_rays, void *stream);

/* batch_mvcnn_planes_voxels_with_ray_marching
 * (mvcnn_with_ray_marching_and_voxels_mapping.py:56-111): front end only; S_new [n][M]. */
int rn_mvcnn_voxel(const RnParams *p, const int32_t *ray_idxs, const float *features, const float *P,
                   const float *P_inv, const float *centre, const float *voxel_grid,
                   int32_t *ray_voxel_indices, int32_t *ray_voxel_count, float *S_new, int64_t n_rays,
                   void *stream);

/* ..._with_depth (same file :222-313): + arg-max voxel over all M slots -> depth_map [n]. */
int rn_mvcnn_voxel_depth(const RnParams *p, const int32_t *ray_idxs, const float *features, const float *P,
                         const float *P_inv, const float *centre, const float *voxel_grid,
                         int32_t *ray_voxel_indices, int32_t *ray_voxel_count, float *S_new,
                         float *depth_map, int64_t n_rays, void *stream);

/* ---------------------------------------------------------------------------------------
 * Resident pipeline (B200-native state; no reference equivalent -- it replaces the host
 * loop of RayNetForwardPass.forward_pass, forward_pass.py:593-748, which re-runs the
 * front end and bounces messages through host memory on every sweep).
 *
 * Per-ray state kept in HBM between sweeps (max_voxels <= 1536; R = rn_row_stride(max_voxels)):
 *   ray_hdr   uint32 [n][2]  first voxel + step signs      (8 B / ray)
 *   codes     uint8  [n][code_stride]  2 bits per traversed voxel, stored as one (lo, hi) pair
 *             of 32-bit bit planes per 32 voxels; code = the axis stepped along to ENTER the
 *             voxel, 3 = none (0.25 B / voxel).  rn_code_stride(M) = bytes per ray.
 *   count     int32  [n]
 *   lin       int32  [n][R]  bricked accumulator offset of every traversed voxel (4 B / voxel), so
 *             that a sweep gathers / scatter-adds without re-deriving coordinates
 *   s_hat     float32 [n][R]  clip_and_renorm(S_voxel_space[r, :count])  (mrf_np.py:4-8)
 *   msgs      float32 [n][R]
 * The two occupancy accumulators of the resident pipeline are BRICKED (4x4x2-voxel 128-byte
 * lines of four 2x2x2-voxel sectors; rn_brick_elems floats, padding included); convert with
 * rn_grid_to_bricks / rn_bricks_to_grid.
 * ------------------------------------------------------------------------------------- */
int64_t rn_code_stride(int32_t max_voxels);
int64_t rn_row_stride(int32_t max_voxels);      /* floats per ray of s_hat / msgs rows: M rounded up to 128 */
int rn_num_classes(void);                       /* number of ray-length classes (see rn_engine_bin_rays) */
int64_t rn_brick_elems(const RnParams *p);      /* floats of a bricked accumulator for p->grid (-1 on error) */

/* row-major float32 [Gx][Gy][Gz] -> bricked (padding voxels := pad) and back; the way back
 * optionally applies compute_occupancy_probabilities (mrf_np.py:206-240). */
int rn_grid_to_bricks(const RnParams *p, const float *grid, float *bricks, float pad, void *stream);
int rn_bricks_to_grid(const RnParams *p, const float *bricks, float *grid, int apply_sigmoid, void *stream);

/* Front end once per reference image: fills starts/ends (may be NULL), ray_hdr, codes,
 * count, s_hat, lin.  axis_centres: float32 [Gx+Gy+Gz] voxel-centre coordinates per axis
 * (rn_axis_centres extracts them from the reference's voxel_grid table).
 * view_ids (may be NULL): int32 [V], the slot inside `features` of each of the V views of
 * this reference image, so that one resident feature volume [n_feature_slots][H+p+1][W+p+1][F]
 * serves every reference image (the reference re-uploads a re-ordered copy per image,
 * forward_pass.py:622-641).  NULL means slots 0..V-1.
 * starts / ends: float32 [n][3] (required, written); plane_scratch: float32 [n][depth_planes],
 * caller-owned scratch for the per-ray plane distribution (required when feat_dim == 32; the library
 * keeps no hidden per-thread buffers, so calls on different streams never share state). */
int rn_engine_frontend(const RnParams *p, const int32_t *ray_idxs, const float *features,
                       const int32_t *view_ids, int32_t n_feature_slots, const float *P,
                       const float *P_inv, const float *centre, const float *axis_centres, float *starts,
                       float *ends, uint32_t *ray_hdr, uint8_t *codes, int32_t *count, float *plane_scratch,
                       float *s_hat, int32_t *lin, int64_t n_rays, void *stream);

/* SURVEY.md 8(f) row 1 -- the MV-CNN feature extractor in front of the path (models.py:90-111,
 * forward_pass.py:181-198): one 'valid' 3x3 convolution to 32 channels with the inference-mode batch
 * normalisation folded into a per-channel affine map, optional ReLU; channels-last float32.
 *   in      [n_images][height][width][channels_in]        (channels_in = 1, 3 or 32)
 *   weights [3][3][channels_in][32]                       (Keras Conv2D kernel layout)
 *   out     [n_images][height-2][width-2][32] = relu?(scale[c] * conv(in)[c] + shift[c])
 * with scale = gamma / sqrt(var + eps), shift = beta + scale * (bias - mean).  Five calls (the last
 * without ReLU) on views zero-padded by 11 pixels give the [V][H+12][W+12][32] feature volume that
 * rn_engine_similarity / rn_raynet_fp read. */
int rn_conv3x3_bn_relu(const float *in, const float *weights, const float *scale, const float *shift, float *out,
                       int32_t n_images, int32_t height, int32_t width, int32_t channels_in, int32_t relu,
                       void *stream);

/* SURVEY.md 8(f) row 2 -- depth maps -> point cloud with the multi-view consistency check
 * (pointcloud.py:76-245, PointcloudFromDepthMaps[WithConsistency]).  Per pixel (v, u) of image i:
 * point = centre_i + depth_i[v][u] * unit ray through (u, v)  ->  points [n][H][W][3];
 * tau [n][H][W] = max over the image's n_neighbors cameras j of |depth_j[round(proj_j(point))] -
 * |point - centre_j||, +inf when a projection leaves image j, when the pixel lies within `borders`
 * of the image edge or when gt (may be NULL) is 0 there.  The caller keeps the points with
 * tau < consistency_threshold (row-major order = the reference's).  neighbors: int32
 * [n][n_neighbors] (NULL: no consistency check, tau = 0 for kept pixels).  P [n][3][4],
 * P_pinv [n][4][3], centre [n][4] are float64 like the reference's camera matrices. */
int rn_fuse_depth_maps(const float *depth, const float *gt, const double *P, const double *P_pinv, const double *centre,
                       const int32_t *neighbors, int32_t n_images, int32_t height, int32_t width, int32_t n_neighbors,
                       int32_t borders, float *points, float *tau, void *stream);

/* Exact Euclidean nearest-neighbour distance of every query point to a target cloud -- the quantity
 * behind the reference's accuracy / completeness metrics (metrics.py:156-236, an sklearn KD-tree there).
 * The caller bins the targets into a uniform grid: sorted_targets float32 [nt][3] ordered by cell id
 * (x fastest: id = (z * dims[1] + y) * dims[0] + x, cell = floor((p - origin) / cell) clamped),
 * cell_start int32 [dims[0]*dims[1]*dims[2] + 1].  origin, dims: HOST pointers to 3 values.
 * max_rings <= 0: search until found.  out float32 [n_query]. */
int rn_nn_grid_distances(const float *query, int64_t n_query, const float *sorted_targets, const int32_t *cell_start,
                         const float *origin, float cell, const int32_t *dims, int32_t max_rings, float *out, void *stream);

/* The two halves of rn_engine_frontend as separate calls, so that a caller can trace the rays of
 * every reference image (no feature maps needed: sample_in_bbox + DDA -> starts, ends, ray_hdr,
 * codes, count) and bin them while the feature maps are still on their way to the device, and run
 * the similarity + plane->voxel mapping (-> s_hat, lin) afterwards.  starts / ends: float32 [n][3],
 * written by the first call and read by the second.  Replaces the same reference code as
 * rn_engine_frontend (raynet_fp.py:43-120 + forward_pass.py:622-663). */
int rn_engine_trace(const RnParams *p, const int32_t *ray_idxs, const float *P_inv, const float *centre,
                    float *starts, float *ends, uint32_t *ray_hdr, uint8_t *codes, int32_t *count, int64_t n_rays,
                    void *stream);
int rn_engine_similarity(const RnParams *p, const float *features, const int32_t *view_ids, int32_t n_feature_slots,
                         const float *P, const float *axis_centres, const float *starts, const float *ends,
                         const uint32_t *ray_hdr, const uint8_t *codes, const int32_t *count, float *plane_scratch,
                         float *s_hat, int32_t *lin, int64_t n_rays, void *stream);

/* Group the rays by length class (class c = ceil(count / 128) for count >= 2, class 0 = the rays
 * BP skips) so that each class runs with the shared memory its rays need.  order: int32 [n]
 * out, the ray positions class by class; inside a class rays follow an 8x8-pixel tiled
 * enumeration of each image when seg_len (= rays per reference image, column-major ids as the
 * reference enumerates them, p->height pixels per column) allows it, else ray order.
 * class_scratch: uint64 [2 * rn_num_classes()] device scratch; on return its first
 * rn_num_classes() entries hold the class sizes (the caller prefix-sums them into the
 * class_offsets of rn_engine_bp_iteration). */
int rn_engine_bin_rays(const RnParams *p, const int32_t *count, int64_t n_rays, int64_t seg_len, int32_t *order,
                       uint64_t *class_scratch, void *stream);

/* One BP sweep over resident state; msgs updated in place, acc_out += messages (bricked grids).
 * order / class_offsets (HOST int64 [rn_num_classes() + 1], may both be NULL): the binning of
 * rn_engine_bin_rays.  first_sweep != 0: messages are taken as all-zero and not read
 * (mrf_np.py:275); first_sweep == 2 additionally promises that acc_in holds ONE value everywhere (the
 * prior right after initialisation, mrf_np.py:285-292), so the sweep does not gather it.  max_count: upper bound on count[] (used when there is no binning). */
int rn_engine_bp_iteration(const RnParams *p, const int32_t *lin,
                           const int32_t *count, const float *s_hat, float *msgs, const float *acc_in,
                           float *acc_out, const int32_t *order, const int64_t *class_offsets,
                           int32_t first_sweep, int32_t max_count, int64_t n_rays, void *stream);

/* Depth pass over resident state, all reference images in one launch:
 * depth_map[r] = |centre(voxel argmax_i o_i cp_i s_i) - C_image(r)|.  centres: float32 [n_seg][4];
 * seg_starts: device int64 [n_seg + 1], first ray of every image (ignored when n_seg == 1).
 * S_new (may be NULL): float32 [n][R], the normalised depth distribution (parity tests). */
int rn_engine_depth(const RnParams *p, const int32_t *lin, const int32_t *count,
                    const float *s_hat, const float *msgs, const float *acc, const float *axis_centres,
                    const float *centres, const int64_t *seg_starts, int32_t n_seg, float *depth_map,
                    float *S_new, int64_t n_rays, void *stream);

/* Expand resident state into the reference's dense buffers (parity tests / debugging):
 * ray_voxel_indices int32 [n][M][3] (zero beyond count). */
int rn_engine_expand_indices(const RnParams *p, const uint32_t *ray_hdr, const uint8_t *codes,
                             const int32_t *count, int32_t *ray_voxel_indices, int64_t n_rays, void *stream);

/* axis_centres[0:Gx] = voxel_grid[i][0][0][0], [Gx:Gx+Gy] = voxel_grid[0][j][0][1],
 * [Gx+Gy:] = voxel_grid[0][0][k][2]  (the table of get_voxel_grid is separable,
 * utils/generic_utils.py:104-110). */
int rn_axis_centres(const RnParams *p, const float *voxel_grid, float *axis_centres, void *stream);

/* acc[i] = prior + partial[i]: the epilogue after the NCCL all-reduce of the per-rank
 * partial accumulators (multi-GPU path). */
int rn_add_prior(float *acc, float prior, int64_t n, void *stream);

/* max over count[0:n] written to *out_max (device int32). */
int rn_max_count(const int32_t *count, int64_t n, int32_t *out_max, void *stream);

/* ---------------------------------------------------------------------------------------
 * Parity mode -- float64 accumulators.  mrf_np.py as it executes under NumPy >= 2
 * keeps both accumulators in float64 (mrf_np.py:285-292), which makes the occupancy-to-ray chain
 * float64 while messages stay float32.  These entry points follow that flavour statement for
 * statement (RED.ADD.F64 scatter-add; float64 o / cp / prefix / suffix; float32 pos, neg, p, log),
 * so that I sweeps can be gated END TO END at 1e-5 against the reference.  Same buffers as their
 * float32 twins except that the accumulators are float64; multi-GPU: all-reduce the float64 grid.
 * ------------------------------------------------------------------------------------- */
int rn_fill_f64(double *dst, double value, int64_t n, void *stream);
int rn_occupancy_f64(const double *acc, float *out, int64_t n, void *stream);      /* mrf_np.py:206-240 */
int rn_grid_to_bricks_f64(const RnParams *p, const double *grid, double *bricks, double pad, void *stream);
/* bricks -> row-major float64 grid (may be NULL) and / or float32 sigmoid(acc) (may be NULL) */
int rn_bricks_to_grid_f64(const RnParams *p, const double *bricks, double *grid, float *occupancy, void *stream);
/* rn_bp_iteration / rn_depth_estimate on the reference's buffers (mrf_cuda.py:37-122), accumulators
 * float64 row-major [Gx][Gy][Gz]; S is clipped / renormalised on the fly exactly as mrf_np.py:4-8. */
int rn_bp_iteration_f64(const RnParams *p, const float *S, const int32_t *ray_voxel_indices,
                        const int32_t *ray_voxel_count, const double *acc_in, float *msgs, double *acc_out,
                        int64_t n_rays, void *stream);
int rn_depth_estimate_f64(const RnParams *p, const float *S, const int32_t *ray_voxel_indices,
                          const int32_t *ray_voxel_count, const double *acc, const float *msgs, float *S_new,
                          int64_t n_rays, void *stream);
/* rn_engine_bp_iteration / rn_engine_depth on the resident state, accumulators float64 bricked. */
int rn_engine_bp_iteration_f64(const RnParams *p, const int32_t *lin, const int32_t *count, const float *s_hat,
                               float *msgs, const double *acc_in, double *acc_out, int32_t first_sweep,
                               int64_t n_rays, void *stream);
int rn_engine_depth_f64(const RnParams *p, const int32_t *lin, const int32_t *count, const float *s_hat,
                        const float *msgs, const double *acc, const float *axis_centres, const float *centres,
                        const int64_t *seg_starts, int32_t n_seg, float *depth_map, float *S_new, int64_t n_rays,
                        void *stream);

/* ---------------------------------------------------------------------------------------
 * SURVEY.md 8(f) row 3 -- backward pass through the unrolled BP (training).  The reference lets TensorFlow
 * differentiate its graph (tf_implementations/forward_backward_pass.py:128-248, mrf/mrf_tf.py:60-271); here
 * the adjoint of each stage is a kernel over the same buffers (voxel lists int32 [n][M][3], rows float32
 * [n][M], accumulators float32 [Gx][Gy][Gz]).  The caller keeps the inputs of every forward sweep (messages,
 * accumulator) as checkpoints and walks them backwards; raynet_b200/training.py does exactly that.
 * scratch: caller-owned float64 scratch, any size >= rn_backward_scratch_bytes(p, 1); the rays are processed
 * in chunks that fit.  Gradient buffers marked += are accumulated into (zero them first).
 * ------------------------------------------------------------------------------------- */
int64_t rn_backward_scratch_bytes(const RnParams *p, int64_t n_rays);
/* adjoint of one sweep (rn_bp_iteration; mrf_tf.py:60-143): msg_in may be NULL (first sweep, all zero);
 * g_msg_out (direct gradient w.r.t. the sweep's output messages) and g_acc_next (gradient w.r.t. the
 * accumulator prior + sum of those messages) may be NULL; g_s += gradient w.r.t. S_norm =
 * clip_and_renorm(S); g_msg_in = gradient w.r.t. msg_in (may alias g_msg_out); g_acc_in += (atomic). */
int rn_bp_sweep_backward(const RnParams *p, const float *S, const int32_t *ray_voxel_indices, const int32_t *ray_voxel_count,
                         const float *acc_in, const float *msg_in, const float *g_msg_out, const float *g_acc_next,
                         float *g_s, float *g_msg_in, float *g_acc_in, double *scratch, int64_t scratch_bytes,
                         int64_t n_rays, void *stream);
/* adjoint of rn_depth_estimate (mrf_tf.py:146-173): g_S_new [n][M] in; g_s +=, g_msgs =, g_acc += out. */
int rn_depth_estimate_backward(const RnParams *p, const float *S, const int32_t *ray_voxel_indices,
                               const int32_t *ray_voxel_count, const float *acc, const float *msgs, const float *g_S_new,
                               float *g_s, float *g_msgs, float *g_acc, double *scratch, int64_t scratch_bytes,
                               int64_t n_rays, void *stream);
/* adjoint of [clip_and_renorm (mrf_tf.py:6-15) +] rn_planes_to_voxels (planes_voxels_mapping.cu:6-92,
 * forward_backward_pass.py:76-125) + softmax.  g_in [n][M]: the gradient w.r.t. S_norm, or, with
 * g_is_wrt_S_voxel_space != 0, w.r.t. S_voxel_space (the clip_and_renorm adjoint already applied, e.g. by
 * rn_clip_renorm_backward).  Out: g_S_voxel_space [n][M] (may be NULL), g_S_planes [n][D], g_scores [n][D]
 * (may be NULL; the gradient w.r.t. the softmax input). */
int rn_planes_to_voxels_backward(const RnParams *p, const float *voxel_grid, const int32_t *ray_voxel_indices,
                                 const int32_t *ray_voxel_count, const float *starts, const float *ends,
                                 const float *S_planes, const float *g_in, int32_t g_is_wrt_S_voxel_space,
                                 float *g_S_voxel_space, float *g_S_planes, float *g_scores, int64_t n_rays, void *stream);
/* adjoint of clip_and_renorm alone: g_s_norm -> gradient w.r.t. the raw rows S. */
int rn_clip_renorm_backward(const RnParams *p, const float *S, const int32_t *ray_voxel_count, const float *g_s_norm,
                            float *g_S, int64_t n_rays, void *stream);
/* losses of tf_implementations/loss_functions.py:4-35 per ray + gradient w.r.t. y_pred times `scale`:
 * kind 0 emd, 1 squared_emd, 2 expected_squared_error (needs voxel lists, voxel_grid, camera centres [n][4]). */
int rn_depth_loss(const RnParams *p, int32_t kind, const float *y_true, const float *y_pred,
                  const int32_t *ray_voxel_indices, const float *voxel_grid, const float *camera_centres, float *loss,
                  float *g_pred, float scale, int64_t n_rays, void *stream);

/* ---------------------------------------------------------------------------------------
 * The exchange step of the multi-GPU path as one kernel over NVLink peer memory: every rank
 * calls it on its own stream after its sweep; result[p][i] = prior + sum_q partial[q][i] lands in EVERY rank's
 * result buffer (slice `rank` of the grid is reduced and broadcast by this rank).  peer_partials / peer_results
 * / peer_flags: HOST arrays of `world` device addresses -- the same buffer as mapped from every rank (peer
 * mappings, e.g. torch.distributed._symmetric_memory); flags: uint32 [n_ctas][world] per rank, zeroed once;
 * epoch: 2 x the number of earlier calls on these flags (the same on all ranks); n_ctas <= SM count (all CTAs
 * of all ranks must be resident: a CTA waits for its twin on every peer); n: float32 elements, multiple of 4.
 * zero_next_partial (may be NULL): a LOCAL buffer of n floats cleared by the same kernel -- the partial of the next
 * sweep when the caller double-buffers its partials (nobody reads it between two exchanges), which removes the
 * fill launch between sweeps.  Replaces ncclAllReduce + the prior epilogue + the fill; the reference has no
 * multi-GPU code.
 * ------------------------------------------------------------------------------------- */
int rn_peer_allreduce_f32(const uint64_t *peer_partials, const uint64_t *peer_results, const uint64_t *peer_flags,
                          float *zero_next_partial, int32_t rank, int32_t world, int32_t n_ctas, uint32_t epoch,
                          float prior, int64_t n, void *stream);

/* The same exchange with the sum and the broadcast done inside the NVSwitch (NVLS): mc_partial / mc_result are the
 * MULTICAST device addresses of the partial and result buffers (the multicast object bound to every rank's copy, e.g.
 * the multicast_ptr of a torch symmetric-memory handle); one multimem.ld_reduce returns the sum over all ranks' partials,
 * one multimem.st writes all ranks' results.  Everything else (flags, epoch, n_ctas, zero_next_partial, prior) as
 * rn_peer_allreduce_f32.  The switch's addition order is unspecified: results agree with rn_peer_allreduce_f32 to
 * float32 rounding, not bit for bit. */
int rn_peer_allreduce_mc_f32(const float *mc_partial, float *mc_result, const uint64_t *peer_flags, float *zero_next_partial,
                             int32_t rank, int32_t world, int32_t n_ctas, uint32_t epoch, float prior, int64_t n, void *stream);

/* ---------------------------------------------------------------------------------------
 * Plane->voxel mapping fused into the first sweep.  rn_engine_similarity = rn_engine_plane_scores (a2: plane-sweep
 * similarity + softmax -> S_planes float32 [n][depth_planes], feat_dim == 32 only) followed by rn_engine_map_planes
 * (a4 + clip_and_renorm -> s_hat, lin).  A caller that is about to run the FIRST sweep straight after a reset (uniform
 * accumulator, no messages) can skip rn_engine_map_planes: rn_engine_first_sweep_mapped builds the rows of every ray
 * inside the sweep kernel (identical arithmetic), uses them in place, writes them once for the later sweeps, and
 * also maps the rays BP skips.  acc_in: any address holding the accumulator's uniform value (the prior).
 * Replaces planes_voxels_mapping.cu:94-118 + mrf_bp.cu:180-204 for that sweep.
 * ------------------------------------------------------------------------------------- */
int rn_engine_plane_scores(const RnParams *p, const float *features, const int32_t *view_ids, int32_t n_feature_slots,
                           const float *P, const float *starts, const float *ends, float *S_planes, int64_t n_rays,
                           void *stream);

/* rn_engine_plane_scores with an explicit plane schedule.  planes_per_pass: 0 = every ray scores all its depth planes
 * in one pass; k > 0 = the planes are swept in blocks of k, each block over all n_rays (the feature vectors a block
 * touches stay L2-resident when the views' feature maps are many times the L2), the softmax once at the end -- the
 * result is bit-identical; < 0 = chosen by the library, which is what rn_engine_plane_scores does (currently always
 * the single pass: the blocked sweep measured slower on the 2 GB feature maps of BASELINE.json configs[4]). */
int rn_engine_plane_scores_passes(const RnParams *p, const float *features, const int32_t *view_ids, int32_t n_feature_slots,
                                  const float *P, const float *starts, const float *ends, float *S_planes, int64_t n_rays,
                                  int32_t planes_per_pass, void *stream);
int rn_engine_map_planes(const RnParams *p, const float *axis_centres, const float *starts, const float *ends,
                         const uint32_t *ray_hdr, const uint8_t *codes, const int32_t *count, const float *S_planes,
                         float *s_hat, int32_t *lin, int64_t n_rays, void *stream);
int rn_engine_first_sweep_mapped(const RnParams *p, const float *axis_centres, const float *starts, const float *ends,
                                 const uint32_t *ray_hdr, const uint8_t *codes, const int32_t *count,
                                 const float *S_planes, int32_t *lin, float *s_hat, float *msgs, const float *acc_in,
                                 float *acc_out, const int32_t *order, const int64_t *class_offsets, int64_t n_rays,
                                 void *stream);

/* ---------------------------------------------------------------------------------------
 * MV-CNN on the tensor cores (csrc/rn_cnn_tc.cuh): the 32 -> 32 channel layers as an implicit-GEMM convolution on
 * tcgen05.mma kind::tf32 with 3 x TF32 split products (float32-level accuracy).  Activations travel between the
 * layers as exact pairs x = hi + lo (hi: low 13 mantissa bits cleared); w_cat float32 [9][64][32]: per tap the rows
 * 0..31 hold the hi part of W[tap][cout][cin], the rows 32..63 its lo part.  rn_conv3x3_bn_relu_split = the
 * CUDA-core layer (rn_conv3x3_bn_relu, e.g. the 3 -> 32 first layer) with a hi / lo epilogue; rn_conv3x3_bn_relu_tc
 * with out_lo == NULL writes the plain float32 result (last layer).  Replaces models.py:90-111 / model.predict
 * (forward_pass.py:622-624) like rn_conv3x3_bn_relu.
 * ------------------------------------------------------------------------------------- */
int rn_conv3x3_bn_relu_split(const float *in, const float *weights, const float *scale, const float *shift, float *out_hi,
                             float *out_lo, int32_t n_images, int32_t height, int32_t width, int32_t channels_in,
                             int32_t relu, void *stream);
int rn_conv3x3_bn_relu_tc(const float *in_hi, const float *in_lo, const float *w_cat, const float *scale, const float *shift,
                          float *out_hi, float *out_lo, int32_t n_images, int32_t height, int32_t width, int32_t relu,
                          void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RAYNET_B200_H */
