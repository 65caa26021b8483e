"""CPU restatement of the reference's depth-map fusion (test infrastructure only).

Follows raynet/pointcloud.py:93-148 (kept pixels, back-projection along the normalised ray),
:177-186 (nearest cameras), :208-240 (consistency: tau = max disagreement, +inf outside a neighbour)
with common/image.py:242-258 (rays) and utils/geometry.py:9-35 (project).  Pinned against an
EXECUTION of those reference functions: tests/golden/pointcloud_golden.npz, produced by
tests/golden/make_pointcloud_golden.py (tests/test_oracle_pinning.py).
"""
import numpy as np


def neighbors(centres, n_neighbors):
    a = np.asarray(centres).T                       # (4, n), pointcloud.py:178-184
    distances = 2 * (a * a).sum(axis=0) - 2 * (a.T.dot(a))
    return distances.argsort()[:, 1:n_neighbors + 1]


def fuse(depth, gt, P, P_pinv, centre, borders, consistency_threshold=None, n_neighbors=0):
    """depth, gt [n,H,W]; P [n,3,4]; P_pinv [n,4,3]; centre [n,4] -> points (3, N) in the reference's order."""
    depth = np.asarray(depth)
    n, H, W = depth.shape
    nb = neighbors(centre, n_neighbors) if consistency_threshold is not None else None
    out = []
    for i in range(n):
        u, v = np.meshgrid(np.arange(W), np.arange(H))                       # [H, W] grids, row-major = reference order
        keep = np.zeros((H, W), bool)
        keep[borders:H - borders, borders:W - borders] = True
        if gt is not None:
            keep &= np.asarray(gt[i]) != 0
        uu, vv = u[keep].astype(np.float64), v[keep].astype(np.float64)
        pix = np.stack([uu, vv, np.ones_like(uu)])
        rays = np.asarray(P_pinv[i], np.float64).dot(pix)
        rays = rays / rays[-1:]
        c = np.asarray(centre[i], np.float64).reshape(4, 1)
        d = rays - c
        pts = c + depth[i][keep].astype(np.float64)[None, :] * d / np.sqrt((d ** 2).sum(axis=0, keepdims=True))
        if nb is not None:
            tau = np.zeros(pts.shape[1])
            for j in nb[i]:
                q = np.asarray(P[j], np.float64).dot(pts)
                x = np.round(q[0] / q[2]).astype(np.int64)
                y = np.round(q[1] / q[2]).astype(np.int64)
                valid = (0 <= x) & (x < W) & (0 <= y) & (y < H)
                x[~valid] = 0
                y[~valid] = 0
                pred = depth[j][y, x].astype(np.float64)
                dist = np.sqrt(((pts - np.asarray(centre[j], np.float64).reshape(4, 1)) ** 2).sum(axis=0))
                tau = np.maximum(tau, np.abs(pred - dist))
                tau[~valid] = np.inf
            pts = pts[:, tau < consistency_threshold]
        out.append(pts[:3])
    return np.hstack(out)
