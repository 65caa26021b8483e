"""Pin the CPU oracle (oracle/rn_oracle.c) against the reference BEFORE trusting it.

Three tiers:
  1. the reference's own golden vectors / scenarios (tests/test_ray_marching.py,
     tests/test_mrf.py, tests/test_planes_voxels_mapping.py of the reference), replayed;
  2. committed fixtures produced by executing the reference (tests/golden/make_golden.py);
  3. live comparison with the reference's own code compiled into oracle/_ref/ (skipped if
     those binaries are absent).
CPU only; runs in the default (not gpu) suite.
"""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _traverse(oracle, bbox, grid, start, end, M):
    idx, cnt = oracle.voxel_traversal(np.array(bbox, np.float32), np.array(grid, np.int32),
                                      np.array(start, np.float32), np.array(end, np.float32), M)
    return idx[0], int(cnt[0])


# ------------------------------------------------------------------ tier 1: reference tests
def test_dda_reference_test_2d(oracle):
    """tests/test_ray_marching.py:20-54 of the reference."""
    bbox, grid = [3, 3, 0, 6, 6, 1], [3, 3, 1]
    v, n = _traverse(oracle, bbox, grid, [3., 4.1, 0.5], [6., 4.9, 0.5], 10)
    assert n == 3
    assert np.all(v[:3, 1] == 1) and np.all(v[:3, 0] == np.arange(3))
    _, n = _traverse(oracle, bbox, grid, [4., 6., 0.5], [6., 5., 0.5], 10)
    assert n == 2
    _, n = _traverse(oracle, bbox, grid, [3., 3., 0.5], [6., 6., 0.5], 10)
    assert n == 5
    _, n = _traverse(oracle, bbox, grid, [6., 6., 0.5], [3., 3., 0.5], 10)
    assert n == 5


def test_dda_reference_golden_list(oracle):
    """tests/test_ray_marching.py:56-77: the exact 9-voxel list."""
    v, n = _traverse(oracle, [0, 0, 0, 6, 6, 1], [6, 6, 1], [0., 3.5, 0.5], [6., 0.5, 0.5], 10)
    assert n == 9
    expect = np.array([[0, 3, 0], [0, 2, 0], [1, 2, 0], [2, 2, 0], [2, 1, 0], [3, 1, 0], [4, 1, 0], [4, 0, 0],
                       [5, 0, 0], [0, 0, 0]])
    assert np.array_equal(v, expect)


def test_dda_reference_test_3d(oracle):
    """tests/test_ray_marching.py:92-102."""
    _, n = _traverse(oracle, [-3., -3., -0.5, 3., 3., 2.], [32, 32, 10], [-1.40056884, -1.34645462, 2.],
                     [-2.30040455, 3., -0.37297964], 100)
    assert 0 < n < 50


def _mrf_scenario(oracle, rays, S_rows, M, acc_f64=False, iters=3):
    bbox = np.array([0, 0, 0, 6, 6, 1], np.float32)
    grid = np.array([6, 6, 1], np.int32)
    starts = np.array([r[0] for r in rays], np.float32)
    ends = np.array([r[1] for r in rays], np.float32)
    idx, cnt = oracle.voxel_traversal(bbox, grid, starts, ends, M)
    S = np.zeros((len(rays), M), np.float32)
    for i, row in enumerate(S_rows):
        S[i, :len(row)] = row
    acc, msgs = oracle.belief_propagation(S, idx, cnt, grid, gamma=0.05, bp_iterations=iters, acc_f64=acc_f64)
    return S, idx, cnt, grid, acc, msgs


RAY1 = ([0., 3.5, .5], [6., .5, .5])
S_PEAK = [0.075, 0.075, 0.075, 0.4, 0.075, 0.075, 0.075, 0.075, 0.075, 0.0]


@pytest.mark.parametrize("acc_f64", [False, True])
def test_mrf_reference_single_ray(oracle, acc_f64):
    """tests/test_mrf.py:36-76: arg-max voxel is (2, 2)."""
    _, _, _, _, acc, _ = _mrf_scenario(oracle, [RAY1], [S_PEAK], 10, acc_f64)
    occ = oracle.occupancy(acc)
    mx = np.where(occ == occ.max())
    assert mx[0][0] == 2 and mx[1][0] == 2


@pytest.mark.parametrize("acc_f64", [False, True])
def test_mrf_reference_two_rays(oracle, acc_f64):
    """tests/test_mrf.py:85-144 and :153-215."""
    _, _, _, _, acc, _ = _mrf_scenario(oracle, [RAY1, ([6., 5.5, .5], [0., 2.5, .5])], [S_PEAK, S_PEAK], 10, acc_f64)
    occ = oracle.occupancy(acc).T
    assert max(occ[0, 4, 3], occ[0, 2, 2]) >= occ.max() - 0
    s2 = [0.07, 0.07, 0.185, 0.07, 0.07, 0.07, 0.185, 0.07, 0.07, 0.07, 0.07]
    _, _, _, _, acc, _ = _mrf_scenario(oracle, [RAY1, ([6., 5.5, .5], [0., .5, .5])], [S_PEAK + [0.0], s2], 11, acc_f64)
    occ = oracle.occupancy(acc).T
    assert occ[0, 2, 2] >= occ.max()


@pytest.mark.parametrize("acc_f64", [False, True])
def test_mrf_reference_three_rays(oracle, acc_f64):
    """tests/test_mrf.py:224-304: ordering of the top-3 voxels."""
    s2 = [0.45, 0.0875, 0.2, 0.0875, 0.0875, 0.0875, 0, 0, 0, 0, 0]
    s3 = [0.07, 0.07, 0.185, 0.07, 0.07, 0.07, 0.185, 0.07, 0.07, 0.07, 0.07]
    rays = [RAY1, ([0., 2.5, .5], [6., 2.5, .5]), ([6., 5.5, .5], [0., .5, .5])]
    _, _, _, _, acc, _ = _mrf_scenario(oracle, rays, [S_PEAK + [0.0], s2, s3], 11, acc_f64)
    occ = oracle.occupancy(acc).T[0]
    order = np.argsort(-occ.ravel())
    top = [np.unravel_index(o, occ.shape) for o in order[:3]]
    assert top[0] == (2, 2) and top[1] == (2, 0) and top[2] == (4, 4)


@pytest.mark.parametrize("acc_f64", [False, True])
def test_mrf_reference_conflict_and_depth(oracle, acc_f64):
    """tests/test_mrf.py:313-349 and :358-416."""
    M = 11
    rows = [np.zeros(M, np.float32), np.zeros(M, np.float32)]
    rows[0][2] = 0.5
    rows[0][6] = 0.5
    rows[1][4] = 1.0
    rays = [RAY1, ([0., 1.5, .5], [4.5, 6., .5])]
    S, idx, cnt, grid, acc, msgs = _mrf_scenario(oracle, rays, rows, M, acc_f64)
    occ = oracle.occupancy(acc).T
    assert occ[0, 0, 2] < 0.1
    S_new = oracle.depth_distribution(S, idx, cnt, grid, acc, msgs, acc_f64=acc_f64)
    assert S_new[0, 2] < 0.5 and S_new[0, 6] > 0.9 and S_new[1, 4] > 0.9


# ------------------------------------------------------------------ tier 2: committed fixtures
def test_dda_matches_reference_fixture(oracle, golden):
    """Bit-exact voxel lists vs the reference's Cython DDA on 800 random rays, 3 grids."""
    for ci in range(3):
        bbox, grid = golden["dda%d_bbox" % ci], golden["dda%d_grid" % ci]
        ref_idx, ref_cnt = golden["dda%d_idx" % ci], golden["dda%d_cnt" % ci]
        idx, cnt = oracle.voxel_traversal(bbox, grid, golden["dda%d_starts" % ci], golden["dda%d_ends" % ci],
                                          ref_idx.shape[1])
        assert np.array_equal(cnt, ref_cnt)
        assert np.array_equal(idx, ref_idx)
        assert ref_cnt.max() > 10


@pytest.mark.parametrize("iters", [1, 3])
def test_bp_matches_reference_fixture(oracle, golden, iters):
    """C restatement (NumPy>=2 precision: f64 accumulators) vs mrf_np executed in the build
    container.  Tolerances on probabilities, as DESIGN.md states."""
    assert str(golden["numpy_version"]).startswith("2."), "fixture was generated under NumPy >= 2"
    S, idx, cnt, grid = golden["bp_S"], golden["bp_idx"], golden["bp_cnt"], golden["bp_grid"]
    acc, msgs = oracle.belief_propagation(S, idx, cnt, grid, gamma=0.05, bp_iterations=iters, acc_f64=True)
    ref_acc, ref_msgs = golden["bp_acc_it%d" % iters], golden["bp_msgs_it%d" % iters]
    assert ref_acc.dtype == np.float64
    occ, ref_occ = oracle.occupancy(acc), golden["bp_occ_it%d" % iters]
    assert np.abs(occ - ref_occ).max() < 2e-6
    sig = lambda x: 1.0 / (1.0 + np.exp(-x.astype(np.float64)))
    assert np.abs(sig(msgs) - sig(ref_msgs)).max() < 2e-6
    S_new = oracle.depth_distribution(S, idx, cnt, grid, ref_acc, ref_msgs, acc_f64=True)
    assert np.abs(S_new - golden["bp_Snew_it%d" % iters]).max() < 2e-6
    # the f32-accumulator flavour (NumPy < 2, and what the CUDA kernels compute in) stays close
    acc32, _ = oracle.belief_propagation(S, idx, cnt, grid, gamma=0.05, bp_iterations=iters, acc_f64=False)
    assert np.abs(oracle.occupancy(acc32) - ref_occ).max() < (1e-5 if iters == 1 else 1e-4)


def test_planes_voxels_matches_reference_fixture(oracle, golden):
    """CUDA-flavoured interpolation (planes_voxels_mapping.cu) vs the reference's numpy li / li_2
    -- the cross-check of tests/test_planes_voxels_mapping.py:61-78 (np.allclose defaults)."""
    T, C, _ = golden["pv_vox"].shape
    D = golden["pv_s"].shape[1]
    for t in range(T):
        vox, pts, s = golden["pv_vox"][t], golden["pv_pts"][t], golden["pv_s"][t]
        # the restatement reads voxel centres from a table through integer indices: build a
        # 1 x 1 x C "grid" whose centres are the C test voxels, sorted along the ray like a DDA list
        ray = pts[:3, -1] - pts[:3, 0]
        order = np.argsort(((vox - pts[:3, 0]) * ray).sum(axis=1))
        table = vox[order].astype(np.float32).reshape(1, 1, C, 3)
        idx = np.zeros((1, C, 3), np.int32)
        idx[0, :, 2] = np.arange(C)
        out = oracle.planes_voxels_mapping(table, [1, 1, C], idx, np.array([C], np.int32),
                                           pts[:3, 0].astype(np.float32).reshape(1, 3),
                                           pts[:3, -1].astype(np.float32).reshape(1, 3),
                                           s.astype(np.float32).reshape(1, D), C)[0]
        assert np.allclose(out, golden["pv_li"][t][order], rtol=1e-4, atol=1e-6)
        assert np.allclose(out, golden["pv_li2"][t][order], rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------ tier 3: live reference binaries
def test_dda_bit_exact_vs_live_cython(oracle, ref_mods):
    if "ray_tracing" not in ref_mods:
        pytest.skip("oracle/_ref/ray_tracing*.so not built")
    rt = ref_mods["ray_tracing"]
    rng = np.random.RandomState(7)
    bbox = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    for G, M in ((32, 96), (128, 384)):
        grid = np.array([G, G, G], np.int32)
        n = 300
        # rays from an exterior camera through the box, like the synthetic rig
        C = np.array([2.6, 0.4, 1.5])
        targets = rng.rand(n, 3) * 2 - 1
        d = targets - C
        t0 = np.full(n, -np.inf); t1 = np.full(n, np.inf)
        for a in range(3):
            ta = (-1 - C[a]) / d[:, a]; tb = (1 - C[a]) / d[:, a]
            t0 = np.maximum(t0, np.minimum(ta, tb)); t1 = np.minimum(t1, np.maximum(ta, tb))
        starts = (C + t0[:, None] * d).astype(np.float32)
        ends = (C + t1[:, None] * d).astype(np.float32)
        idx, cnt = oracle.voxel_traversal(bbox, grid, starts, ends, M)
        for r in range(n):
            ref = np.zeros((M, 3), np.int32)
            c = rt.voxel_traversal(bbox, grid, ref, starts[r], ends[r])
            assert c == cnt[r]
            assert np.array_equal(ref, idx[r])


def test_bp_vs_live_mrf_np(oracle, ref_mods, golden):
    if "ref_mrf_np" not in ref_mods:
        pytest.skip("oracle/_ref/ref_mrf_np*.so not built")
    mrf = ref_mods["ref_mrf_np"]
    S, idx, cnt, grid = golden["bp_S"][:120], golden["bp_idx"][:120], golden["bp_cnt"][:120], golden["bp_grid"]
    msgs = np.zeros_like(S)
    ref_acc, ref_msgs = mrf.belief_propagation(S, idx, cnt, msgs, grid, gamma=0.05, bp_iterations=2)
    f64 = np.asarray(ref_acc).dtype == np.float64
    acc, m = oracle.belief_propagation(S, idx, cnt, grid, gamma=0.05, bp_iterations=2, acc_f64=f64)
    assert np.abs(oracle.occupancy(acc) - mrf.compute_occupancy_probabilities(ref_acc)).max() < 2e-6


def test_voxel_grid_table_is_separable(oracle):
    """The product keeps only three axis slices of the reference's voxel-centre table
    (utils/generic_utils.py:104-110); check the table really is separable and that the oracle's
    table equals the numpy statements the reference executes (as they run under this NumPy)."""
    from raynet_b200.synth import get_voxel_grid
    for bbox, grid in (([-1, -1, -1, 1, 1, 1], [32, 16, 8]), ([-0.7, -0.3, 0.1, 1.5, 2.2, 0.9], [12, 20, 7])):
        vg = get_voxel_grid(bbox, grid)
        assert np.array_equal(vg[0], np.broadcast_to(vg[0, :, :1, :1], vg[0].shape))
        assert np.array_equal(vg[1], np.broadcast_to(vg[1, :1, :, :1], vg[1].shape))
        assert np.array_equal(vg[2], np.broadcast_to(vg[2, :1, :1, :], vg[2].shape))
    vg = get_voxel_grid([-1, -1, -1, 1, 1, 1], [32, 16, 8]).transpose(1, 2, 3, 0)
    assert np.array_equal(vg, oracle.voxel_grid(np.array([-1, -1, -1, 1, 1, 1], np.float32), [32, 16, 8]))


def test_reference_cuda_kernels_compile_for_sm100a():
    """oracle/build_ref_cuda.py: the reference's own .cu templates + the kernel text of raynet_fp.py,
    substituted like PyCUDA's SourceModule would, compile with nvcc for sm_100a and export both fused
    kernels (the GPU-side oracle of tests/test_gpu_ref_cuda.py).  Needs /root/reference and nvcc."""
    import os
    import shutil
    import subprocess
    from oracle import build_ref_cuda
    if not os.path.isdir("/root/reference/raynet") or shutil.which("nvcc") is None:
        pytest.skip("reference tree or nvcc absent")
    assert build_ref_cuda.build()
    for name in build_ref_cuda.CONFIGS:
        cubin = os.path.join(build_ref_cuda.OUT, "raynet_fp_%s.cubin" % name)
        assert os.path.exists(cubin)
    out = subprocess.run(["cuobjdump", "-elf", os.path.join(build_ref_cuda.OUT, "raynet_fp_c1.cubin")],
                         capture_output=True, text=True).stdout
    assert ".text.batch_raynet_fp" in out and ".text.batch_complete_depth_estimation" in out


def test_cnn_oracle_convolution_semantics():
    """oracle/cnn_np.py against an explicit loop: 'valid' cross-correlation, channels-last,
    Keras kernel layout [kh][kw][cin][cout] (models.py:90-111)."""
    from oracle import cnn_np
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 6, 7, 3))
    k = rng.normal(size=(3, 3, 3, 4))
    b = rng.normal(size=4)
    ref = np.zeros((2, 4, 5, 4))
    for n in range(2):
        for y in range(4):
            for xx in range(5):
                for o in range(4):
                    ref[n, y, xx, o] = b[o] + sum(x[n, y + dy, xx + dx, c] * k[dy, dx, c, o]
                                                  for dy in range(3) for dx in range(3) for c in range(3))
    assert np.abs(cnn_np.conv3x3_valid(x, k, b) - ref).max() < 1e-12


def test_cnn_oracle_vs_independent_library_network():
    """The whole oracle network against the same five blocks written with another library's operators in float64
    (torch.nn.functional.conv2d = cross-correlation on channels-first tensors, batch_norm in inference mode with
    eps = 1e-3, relu).  Not the reference (Keras is not installed) -- an independent second implementation of the
    layer semantics models.py:90-111 names; the CNN row stays "parity unpinned"."""
    import torch
    import torch.nn.functional as F
    from oracle import cnn_np
    from raynet_b200.models import SimpleCNN
    m = SimpleCNN.random_init(channels=3, seed=5)
    weights = m.get_weights()
    rng = np.random.default_rng(1)
    x = rng.uniform(0, 1, size=(2, 19, 23, 3))
    ours = cnn_np.simple_cnn_forward(x, weights)
    t = torch.from_numpy(x).permute(0, 3, 1, 2).double()
    n_layers = len(weights) // 6
    for l in range(n_layers):
        k, b, g, be, mu, var = [torch.from_numpy(np.asarray(w)).double() for w in weights[6 * l:6 * l + 6]]
        t = F.conv2d(t, k.permute(3, 2, 0, 1), b)                       # [kh][kw][cin][cout] -> [cout][cin][kh][kw]
        t = F.batch_norm(t, mu, var, g, be, training=False, eps=1e-3)
        if l < n_layers - 1:
            t = F.relu(t)
    theirs = t.permute(0, 2, 3, 1).numpy()
    assert ours.shape == theirs.shape == (2, 9, 13, 32)
    assert np.abs(ours - theirs).max() < 1e-12


GOLDEN_PC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pointcloud_golden.npz")


@pytest.mark.parametrize("borders,thr,nn", [(4, 0.05, 2), (0, 0.02, 3)])
def test_pointcloud_oracle_vs_reference_execution(borders, thr, nn):
    """oracle/pointcloud_np.py against the fixture made by executing raynet/pointcloud.py itself."""
    from oracle import pointcloud_np
    g = np.load(GOLDEN_PC)
    key = "b%d_t%g_n%d" % (borders, thr, nn)
    assert np.array_equal(pointcloud_np.neighbors(g["centre"], nn), g["neigh_" + key])
    plain = pointcloud_np.fuse(g["depth"], g["gt"], g["P"], g["P_pinv"], g["centre"], borders)
    assert plain.shape == g["plain_" + key].shape and np.abs(plain - g["plain_" + key]).max() < 1e-12
    cons = pointcloud_np.fuse(g["depth"], g["gt"], g["P"], g["P_pinv"], g["centre"], borders, thr, nn)
    assert cons.shape == g["cons_" + key].shape and np.abs(cons - g["cons_" + key]).max() < 1e-12


# ----------------------------------------------------------------------------- the autograd oracle
def _training_case(oracle, n_rays=120, seed=3):
    from rig import Case
    c = Case(24, 2, 8, 32, 32, 72, n_rays=n_rays, seed=seed)
    o = oracle.frontend(c.ray_idxs, c.features, c.P, c.P_inv, c.centre, c.vgrid, c.bbox, c.grid, c.M, c.D, c.V, 32,
                        c.H, c.W, 11, want_stages=True)
    rng = np.random.RandomState(seed)
    scores = rng.randn(c.N, c.D) * 2.0
    target = np.zeros((c.N, c.M), np.float32)
    for r in range(c.N):
        L = int(o["cnt"][r])
        if L > 0:
            t = rng.rand(L) ** 4
            target[r, :L] = t / t.sum()
    return c, o, scores, target


def test_autograd_restatement_reproduces_the_c_oracle(oracle):
    """oracle/bp_autograd.py (float64 torch restatement of mrf_tf.py + the plane->voxel interpolation) must
    reproduce rn_oracle.c's forward pass -- that is what pins the reference gradients used by the GPU tests."""
    import torch
    from oracle import bp_autograd as ag
    c, o, scores, target = _training_case(oracle)
    S = torch.softmax(torch.from_numpy(scores), dim=1)
    starts, ends = o["starts"], o["ends"]
    S32 = S.numpy().astype(np.float32)
    S_vox_c = oracle.planes_voxels_mapping(c.vgrid, c.grid, o["idx"], o["cnt"], starts, ends, S32, c.M)
    acc_c, msgs_c = oracle.belief_propagation(S_vox_c, o["idx"], o["cnt"], c.grid, gamma=0.05, bp_iterations=3, acc_f64=True)
    Smrf_c = oracle.depth_distribution(S_vox_c, o["idx"], o["cnt"], c.grid, acc_c, msgs_c, acc_f64=True)
    loss, S_mrf, S_vox = ag.forward_graph(torch.from_numpy(scores), o["idx"], o["cnt"], c.grid, c.vgrid, starts, ends,
                                          np.tile(np.append(c.centre[:3], 1.0), (c.N, 1)), target,
                                          torch.tensor(0.05, dtype=torch.float64), 3, "squared_emd")
    assert np.abs(S_vox.numpy() - S_vox_c).max() <= 2e-6
    assert np.abs(S_mrf.numpy() - Smrf_c).max() <= 1e-5
    assert np.isfinite(float(loss))


def test_autograd_restatement_gradcheck(oracle):
    """float64 finite differences of the restatement itself (torch.autograd.gradcheck) on a tiny problem."""
    import torch
    from oracle import bp_autograd as ag
    c, o, scores, target = _training_case(oracle, n_rays=6, seed=5)
    cam = np.tile(np.append(c.centre[:3], 1.0), (c.N, 1))
    for loss in ("squared_emd", "expected_squared_error"):
        def f(z, g):
            return ag.forward_graph(z, o["idx"], o["cnt"], c.grid, c.vgrid, o["starts"], o["ends"], cam, target, g, 2, loss)[0]
        z = torch.from_numpy(scores).requires_grad_(True)
        g = torch.tensor(0.05, dtype=torch.float64, requires_grad=True)
        assert torch.autograd.gradcheck(f, (z, g), eps=1e-6, atol=1e-7, rtol=1e-4)
