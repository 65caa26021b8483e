"""Generate the golden fixtures of tests/golden/ by EXECUTING THE REFERENCE.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
It imports the reference's own code as compiled by oracle/build_ref.py (Cython DDA from
ray_tracing.pyx; mrf_np.py / planes_voxels_mapping.py through the 2-regex py2->py3 shim) and
stores small input/output vectors.  NumPy version is recorded: under NumPy >= 2 the
reference's accumulators are float64 (np.ones(f32) * np.float64), see oracle/rn_oracle.c.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref, oracle as orc  # noqa: E402


def random_rays_on_bbox(rng, bbox, n):
    """start / end points on (or slightly inside/outside) the faces of the bbox."""
    lo, hi = bbox[:3], bbox[3:]
    starts = np.zeros((n, 3), np.float32)
    ends = np.zeros((n, 3), np.float32)
    for r in range(n):
        for arr in (starts, ends):
            p = lo + rng.rand(3) * (hi - lo)
            a = rng.randint(3)
            p[a] = lo[a] if rng.rand() < 0.5 else hi[a]
            arr[r] = p.astype(np.float32)
    return starts, ends


def main():
    assert build_ref.build(), "reference tree not available"
    mods = orc.ref_modules()
    rt, mrf, pvm = mods["ray_tracing"], mods["ref_mrf_np"], mods["ref_planes_voxels_mapping"]
    rng = np.random.RandomState(1234)
    out = {}

    # ---- DDA: reference Cython on random rays, three grids ---------------------------------
    dda_cases = [
        (np.array([-1, -1, -1, 1, 1, 1], np.float32), np.array([32, 32, 32], np.int32), 96, 300),
        (np.array([-0.7, -0.3, 0.1, 1.5, 2.2, 0.9], np.float32), np.array([48, 40, 12], np.int32), 110, 300),
        (np.array([3, 3, 1, 6, 6, 2], np.float32), np.array([64, 64, 15], np.int32), 256, 200),
    ]
    for ci, (bbox, grid, M, n) in enumerate(dda_cases):
        starts, ends = random_rays_on_bbox(rng, bbox.astype(np.float64), n)
        # a few degenerate ones: axis-parallel, reversed, starting outside
        starts[0] = [bbox[0], (bbox[1] + bbox[4]) / 2, (bbox[2] + bbox[5]) / 2]
        ends[0] = [bbox[3], (bbox[1] + bbox[4]) / 2, (bbox[2] + bbox[5]) / 2]
        starts[1], ends[1] = ends[0].copy(), starts[0].copy()
        starts[2] = bbox[:3] - 1.0
        idx = np.zeros((n, M, 3), np.int32)
        cnt = np.zeros((n,), np.int32)
        for r in range(n):
            cnt[r] = rt.voxel_traversal(bbox, grid, idx[r], starts[r], ends[r])
        out["dda%d_bbox" % ci], out["dda%d_grid" % ci] = bbox, grid
        out["dda%d_starts" % ci], out["dda%d_ends" % ci] = starts, ends
        out["dda%d_idx" % ci], out["dda%d_cnt" % ci] = idx, cnt

    # ---- BP: reference mrf_np on a random 16^3 problem ------------------------------------------
    bbox = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    grid = np.array([16, 16, 16], np.int32)
    N, M = 400, 48
    starts, ends = random_rays_on_bbox(rng, bbox.astype(np.float64), N)
    idx = np.zeros((N, M, 3), np.int32)
    cnt = np.zeros((N,), np.int32)
    for r in range(N):
        cnt[r] = rt.voxel_traversal(bbox, grid, idx[r], starts[r], ends[r])
    S = np.zeros((N, M), np.float32)
    for r in range(N):
        c = cnt[r]
        if c > 0:
            s = rng.rand(c) ** 4 + 1e-3
            S[r, :c] = (s / s.sum()).astype(np.float32)
    for iters in (1, 3):
        msgs = rng.rand(N, M).astype(np.float32)       # overwritten with 0 by the reference
        acc, msgs = mrf.belief_propagation(S, idx, cnt, msgs, grid, gamma=0.05, bp_iterations=iters)
        S_new = mrf.compute_depth_distribution(S, idx, cnt, msgs, acc, np.zeros_like(S))
        out["bp_acc_it%d" % iters] = np.asarray(acc)
        out["bp_msgs_it%d" % iters] = np.asarray(msgs)
        out["bp_Snew_it%d" % iters] = np.asarray(S_new)
        out["bp_occ_it%d" % iters] = np.asarray(mrf.compute_occupancy_probabilities(acc))
    out["bp_grid"], out["bp_S"], out["bp_idx"], out["bp_cnt"] = grid, S, idx, cnt

    # ---- plane -> voxel: reference numpy li / li_2 -----------------------------------------
    C, D, T = 10, 5, 10
    pv_vox = np.zeros((T, C, 3)); pv_pts = np.zeros((T, 4, D)); pv_s = np.zeros((T, D))
    pv_li = np.zeros((T, C)); pv_li2 = np.zeros((T, C))
    for t in range(T):   # the generator of tests/test_planes_voxels_mapping.py:61-78, seeded
        voxels = rng.rand(C, 3)
        ps = rng.rand(4, 1) - 1; ps[-1:] = 1
        pe = rng.rand(4, 1) + 1; pe[-1:] = 1
        points = ps + np.linspace(0, 1, D) * (pe - ps)
        s = rng.rand(D); s /= s.sum()
        pv_vox[t], pv_pts[t], pv_s[t] = voxels, points, s
        pv_li[t] = pvm.single_ray_depth_to_voxels_li(voxels.T, points[:-1], s)
        pv_li2[t] = pvm.single_ray_depth_to_voxels_li_2(voxels.T, points[:-1], s)
    out.update(pv_vox=pv_vox, pv_pts=pv_pts, pv_s=pv_s, pv_li=pv_li, pv_li2=pv_li2)
    out["numpy_version"] = np.array(np.__version__)
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_golden.npz"), "numpy", np.__version__)


if __name__ == "__main__":
    main()
