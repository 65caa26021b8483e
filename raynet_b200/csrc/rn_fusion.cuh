// rn_fusion.cuh -- SURVEY.md 8(f) row 2: depth-map fusion into a point cloud with the multi-view
// consistency check of the reference (pointcloud.py:76-245).
//
// One thread per pixel of every depth map: back-project the pixel (image.py:242-258: P_pinv (u, v, 1)
// dehomogenised), walk `depth` along the normalised ray from the camera centre (pointcloud.py:127-148),
// then re-project the point into each of the image's n nearest cameras, read that camera's predicted
// depth at the rounded pixel (np.round: half to even) and keep tau = the largest disagreement with
// the point's distance to that camera (pointcloud.py:208-240); a projection outside the image makes
// tau infinite.  Pixels inside the border or without ground truth get tau = +inf as well
// (pointcloud.py:93-124), so the host keeps exactly the points with tau < threshold, in the
// reference's order (rows of the cropped map, then columns).  Geometry in double like the NumPy
// reference (camera matrices are float64 there); the work is ~100 flops per pixel and neighbour.
#pragma once

#include "rn_common.cuh"

struct FuseArgs {
    const float *depth;        // [n_img][H][W]
    const float *gt;           // [n_img][H][W] or null: pixels with gt == 0 are dropped
    const double *P;           // [n_img][3][4]
    const double *P_pinv;      // [n_img][4][3]
    const double *centre;      // [n_img][4]
    const int32_t *neighbors;  // [n_img][n_nb] indices into the image list, or null (no consistency check)
    float *points;             // [n_img][H][W][3]
    float *tau;                // [n_img][H][W]
    int n_img, H, W, n_nb, borders;
};

__global__ void __launch_bounds__(256) fuse_depth_kernel(FuseArgs a) {
    const int64_t total = (int64_t)a.n_img * a.H * a.W;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int u = (int)(t % a.W), v = (int)((t / a.W) % a.H), img = (int)(t / ((int64_t)a.W * a.H));
    const double *Pi = a.P_pinv + img * 12, *C = a.centre + img * 4;
    // ray through the pixel: P_pinv (u, v, 1), dehomogenised (utils/geometry.py:25-29)
    double r[4];
#pragma unroll
    for (int i = 0; i < 4; i++) r[i] = Pi[i * 3 + 0] * (double)u + Pi[i * 3 + 1] * (double)v + Pi[i * 3 + 2];
    double dir[3], nrm = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) { dir[i] = r[i] / r[3] - C[i]; nrm += dir[i] * dir[i]; }
    nrm = sqrt(nrm);
    const double d = (double)a.depth[t];
    double X[3];
#pragma unroll
    for (int i = 0; i < 3; i++) X[i] = C[i] + d * dir[i] / nrm;
#pragma unroll
    for (int i = 0; i < 3; i++) a.points[3 * t + i] = (float)X[i];
    const bool kept = v >= a.borders && v < a.H - a.borders && u >= a.borders && u < a.W - a.borders &&
                      (a.gt == nullptr || a.gt[t] != 0.f);
    double tau = kept ? 0.0 : (double)INFINITY;
    if (kept && a.neighbors) {
        for (int k = 0; k < a.n_nb; k++) {
            const int j = a.neighbors[img * a.n_nb + k];
            const double *P = a.P + j * 12, *Cj = a.centre + j * 4;
            double q[3];
#pragma unroll
            for (int i = 0; i < 3; i++) q[i] = P[i * 4 + 0] * X[0] + P[i * 4 + 1] * X[1] + P[i * 4 + 2] * X[2] + P[i * 4 + 3];
            const double px = rint(q[0] / q[2]), py = rint(q[1] / q[2]);     // np.round: half to even
            const bool valid = px >= 0.0 && px < (double)a.W && py >= 0.0 && py < (double)a.H;
            if (!valid) { tau = (double)INFINITY; continue; }    // pointcloud.py:187: invalid in ANY neighbour rejects
            const double pred = (double)a.depth[((int64_t)j * a.H + (int)py) * a.W + (int)px];
            double dist = 0.0;
#pragma unroll
            for (int i = 0; i < 3; i++) dist += (X[i] - Cj[i]) * (X[i] - Cj[i]);
            tau = fmax(tau, fabs(pred - sqrt(dist)));
        }
    }
    a.tau[t] = (float)tau;
}

// =======================================================================================
// Nearest-neighbour distances between two point clouds (pointcloud.py:64-73, metrics.py:156-236:
// accuracy = distance of every predicted point to the ground-truth cloud, completeness the other
// way round; the reference asks an sklearn KD-tree).  Here the target cloud is binned into a
// uniform grid on the host side (torch sort by cell + cell_start offsets); one thread per query
// scans the cells ring by ring (Chebyshev distance r around its own cell) and stops as soon as
// the best squared distance is within (r h)^2: every unscanned point lies in a cell at Chebyshev
// distance >= r + 1 and is therefore further than r h away.  Exact, Euclidean, float32.
// =======================================================================================
struct NnArgs {
    const float *query;        // [nq][3]
    const float *target;       // [nt][3] sorted by cell
    const int32_t *cell_start; // [nx*ny*nz + 1]
    float *out;                // [nq] distance to the nearest target point
    int64_t nq;
    float origin[3];
    float cell;                // edge length h
    int dims[3];
    int max_rings;
};

__global__ void __launch_bounds__(128) nn_grid_kernel(NnArgs a) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nq) return;
    const float q[3] = {a.query[3 * t], a.query[3 * t + 1], a.query[3 * t + 2]};
    int c[3];
#pragma unroll
    for (int i = 0; i < 3; i++) c[i] = min(max((int)floorf((q[i] - a.origin[i]) / a.cell), 0), a.dims[i] - 1);
    float best = INFINITY;
    const int rmax = min(a.max_rings, max(a.dims[0], max(a.dims[1], a.dims[2])));
    for (int r = 0; r <= rmax; r++) {
        const int z0 = max(c[2] - r, 0), z1 = min(c[2] + r, a.dims[2] - 1);
        const int y0 = max(c[1] - r, 0), y1 = min(c[1] + r, a.dims[1] - 1);
        const int x0 = max(c[0] - r, 0), x1 = min(c[0] + r, a.dims[0] - 1);
        for (int z = z0; z <= z1; z++) {
            for (int y = y0; y <= y1; y++) {
                const bool face = (abs(z - c[2]) == r) || (abs(y - c[1]) == r);   // whole x-run is on the shell
                for (int x = x0; x <= x1; x += (face ? 1 : max(x1 - x0, 1))) {    // otherwise only its two ends
                    if (!face && abs(x - c[0]) != r) continue;
                    const int64_t cellid = ((int64_t)z * a.dims[1] + y) * a.dims[0] + x;
                    const int s = a.cell_start[cellid], e = a.cell_start[cellid + 1];
                    for (int k = s; k < e; k++) {
                        const float dx = a.target[3 * k] - q[0], dy = a.target[3 * k + 1] - q[1], dz = a.target[3 * k + 2] - q[2];
                        best = fminf(best, dx * dx + dy * dy + dz * dz);
                    }
                }
            }
        }
        const float reach = (float)r * a.cell;
        if (best <= reach * reach) break;
    }
    a.out[t] = sqrtf(best);
}
