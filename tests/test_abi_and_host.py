"""CPU-side checks: the C-ABI library loads and exports every symbol of include/raynet_b200.h, the
host logic (parameter block, ray sharding, error mapping) behaves, the product never touches
oracle/, and the N>1 path (ray blocks + one SUM all-reduce per sweep) reproduces the
single-process result over a world_size-2 gloo group.  No kernel is launched here.
"""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "raynet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from raynet_b200 import _lib, build
    build.build()                                  # nvcc cross-compiles without a GPU
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "symbol %s declared in the header is not exported" % s
    # and the ctypes table covers the header exactly
    assert sorted(list(_lib.SIGNATURES) + _lib.OTHER_SYMBOLS) == syms
    assert lib.rn_abi_version() == 3
    assert lib.rn_row_stride(96) == 128 and lib.rn_row_stride(768) == 768 and lib.rn_num_classes() == 13
    assert lib.rn_code_stride(768) == 192 and lib.rn_code_stride(96) == 32 and lib.rn_code_stride(1) == 32


def test_params_struct_layout_matches_header():
    from raynet_b200 import _lib
    p = _lib.make_params(768, 64, 9, 32, 512, 512, 11, [-1, -1, -1, 1, 1, 1], [256, 256, 256])
    assert ctypes.sizeof(p) == 4 * 10 + 4 * 6
    assert (p.max_voxels, p.depth_planes, p.n_views, p.feat_dim, p.height, p.width, p.padding) == \
        (768, 64, 9, 32, 512, 512, 11)
    assert list(p.grid) == [256, 256, 256] and list(p.bbox) == [-1, -1, -1, 1, 1, 1]


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from raynet_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.RayNetB200Error):
        _lib.load()


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from raynet_b200.cuda_implementations.utils import device
    with pytest.raises(RuntimeError):
        device()
    from raynet_b200.mrf.bp_inference import get_bp_backend
    with pytest.raises(NotImplementedError):
        get_bp_backend("numpy", None)
    from raynet_b200.ray_marching.ray_marching import get_voxel_traversal_backend
    with pytest.raises(NotImplementedError):
        get_voxel_traversal_backend("cython")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "raynet_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "rn_oracle" not in src and "liboracle" not in src, f


def test_ray_blocks_partition():
    from raynet_b200 import sharding
    for n in (0, 1, 7, 8, 2359296, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [sharding.ray_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    segs = [sharding.image_segments([10, 10, 10], r, 4) for r in range(4)]
    covered = sorted((k, i) for s in segs for (k, a, b) in s for i in range(a, b))
    assert covered == [(k, i) for k in range(3) for i in range(10)]
    assert sharding.images_of_rank(72, 3, 8) == list(range(27, 36))
    # aligned blocks (whole groups of image columns): C3's 9 x 512 x 512 rays over 8 ranks in units of 64 columns
    H, n_img = 512, 9
    for world in (1, 2, 4, 8, 3):
        segs = [sharding.image_segments([H * H] * n_img, r, world, 64 * H) for r in range(world)]
        flat = [(k, a, b) for sg in segs for (k, a, b) in sg]
        assert all(a % (64 * H) == 0 and b % (64 * H) == 0 for (_, a, b) in flat)
        assert sum(b - a for (_, a, b) in flat) == n_img * H * H
        per_rank = [sum(b - a for (_, a, b) in sg) for sg in segs]
        assert max(per_rank) - min(per_rank) <= 64 * H
        for k in range(n_img):      # every image covered exactly once, in order
            pieces = sorted((a, b) for (kk, a, b) in flat if kk == k)
            assert pieces[0][0] == 0 and pieces[-1][1] == H * H
            assert all(pieces[i][1] == pieces[i + 1][0] for i in range(len(pieces) - 1))
    assert sharding.aligned_ray_block(1000, 1, 3, 64) == sharding.ray_block(1000, 1, 3)      # not a multiple: unit 1
    # work-balanced cuts: equal total weight per block up to one unit, monotone, deterministic
    rng = np.random.RandomState(0)
    w = rng.rand(576) ** 3 + np.linspace(0, 2, 576)
    for world in (1, 2, 4, 8, 5):
        b = sharding.balanced_boundaries(w, world)
        assert b[0] == 0 and b[-1] == 576 and len(b) == world + 1 and all(b[i] <= b[i + 1] for i in range(world))
        work = np.array([w[b[i]:b[i + 1]].sum() for i in range(world)])
        assert work.max() - work.min() <= 2 * w.max() + 1e-9
    assert sharding.balanced_boundaries(np.zeros(10), 3) == [0, 4, 7, 10]
    segs = sharding.segments_from_unit_boundaries([64, 64, 64], 8, 5, 19)
    assert segs == [(0, 40, 64), (1, 0, 64), (2, 0, 24)]
    assert sharding.image_segments([100, 64], 0, 2, 64) == sharding.image_segments([100, 64], 0, 2, 1)
    assert sharding.seed_value(0, -2.5) == -2.5 and sharding.seed_value(3, -2.5) == 0.0


def test_generation_parameters_from_options():
    import argparse
    from raynet_b200.common.generation_parameters import GenerationParameters
    ns = argparse.Namespace(patch_shape=(11, 11, 3), depth_planes=32, neighbors=4, grid_shape=(256, 256, 128),
                            maximum_number_of_marched_voxels=650, depth_range=None, step_depth=None, padding=None,
                            initial_gamma_prior=0.05)
    gp = GenerationParameters.from_options(ns)
    assert gp.padding == 11 and gp.max_number_of_marched_voxels == 650 and gp.gamma_mrf == 0.05
    assert gp.depth_planes == 32 and gp.neighbors == 4


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
from oracle import oracle as orc
from raynet_b200 import sharding
from rig import case_c1
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
c = case_c1()
o = orc.frontend(c.ray_idxs, c.features, c.P, c.P_inv, c.centre, c.vgrid, c.bbox, c.grid, c.M, c.D, c.V, 32,
                 c.H, c.W, 11)
prior = np.float32(np.log(0.05) - np.log(1 - 0.05))
a, b = sharding.ray_block(c.N, rank, world)
S, idx, cnt = o["S_vox"][a:b], o["idx"][a:b], o["cnt"][a:b]
acc_prev = np.full(tuple(c.grid), prior, np.float32)
msgs = np.zeros((b - a, c.M), np.float32)
for it in range(3):
    part = np.full(tuple(c.grid), sharding.seed_value(rank, prior), np.float32)
    orc.bp_iteration(S, idx, cnt, c.grid, acc_prev, part, msgs)
    t = torch.from_numpy(part)
    sharding.allreduce_accumulator(t)
    acc_prev = t.numpy().copy()
if rank == 0:
    ref_acc, ref_msgs = orc.belief_propagation(o["S_vox"], o["idx"], o["cnt"], c.grid, gamma=0.05, bp_iterations=3)
    err = np.abs(orc.occupancy(acc_prev) - orc.occupancy(ref_acc)).max()
    merr = np.abs(msgs - ref_msgs[a:b]).max()
    print("RESULT %.3e %.3e" % (err, merr))
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_gloo_sharded_bp_matches_single_process(tmp_path, oracle):
    """world_size 2 over gloo on CPU: each rank owns a contiguous ray block, partial
    accumulators are seeded (prior on rank 0, zero elsewhere) and summed with the product's
    all-reduce helper after every sweep; the oracle stands in for the kernels.  The result must
    equal the single-process oracle run (summation order differs -> tolerance, not equality)."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0].splitlines() if l.startswith("RESULT")][0]
    err, merr = (float(x) for x in line.split()[1:])
    assert err <= 1e-5 and merr <= 1e-3


# ----------------------------------------------------------------------------- on-disk scenes
GOLDEN_SCENE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scene_golden.npz")


def test_adjacent_frames_match_reference_execution():
    """get_adjacent_frames_idxs against 500+ outputs of the reference's own function
    (tests/golden/make_scene_golden.py) and the four cases of the reference's tests/test_scene.py:110-120."""
    from raynet_b200.common.scene import get_adjacent_frames_idxs
    g = np.load(GOLDEN_SCENE)
    for (ref, n_frames, n_adj, skip), want in zip(g["cases"], g["neighbors"]):
        got = get_adjacent_frames_idxs(int(ref), int(n_frames), int(n_adj), int(skip))
        assert np.array_equal(got, want[:n_adj]), (ref, n_frames, n_adj, skip, got, want)
    assert np.array_equal(get_adjacent_frames_idxs(0, 50, 4, 0), [1, 2, 3, 4])
    assert np.array_equal(get_adjacent_frames_idxs(1, 50, 4, 0), [0, 2, 3, 4])
    assert np.array_equal(get_adjacent_frames_idxs(35, 50, 4, 0), [33, 34, 36, 37])
    assert np.array_equal(get_adjacent_frames_idxs(50, 50, 4, 0), [46, 47, 48, 49])
    with pytest.raises(ValueError):
        get_adjacent_frames_idxs(51, 50, 4, 0)


def test_restrepo_scene_on_disk(tmp_path):
    """A scene written in the Restrepo layout reads back: sorted files, K / R / t, bounding box, image
    scaling, neighbours, voxel grid; and, when the reference tree is present, its mock dataset parses to
    the values the reference's own parser produced (tests/golden/scene_golden.npz)."""
    from PIL import Image as PILImage
    from raynet_b200.common.scene import RestrepoScene, parse_scene_info
    from raynet_b200.synth import ring_cameras
    H, W, n = 12, 16, 6
    os.makedirs(tmp_path / "imgs")
    os.makedirs(tmp_path / "cams_krt")
    cams = ring_cameras(n, H, W)
    rng = np.random.RandomState(0)
    pix = rng.randint(0, 256, size=(n, H, W, 3)).astype(np.uint8)
    for k, c in enumerate(cams):
        PILImage.fromarray(pix[k]).save(str(tmp_path / "imgs" / ("frame%05d.png" % k)))
        with open(str(tmp_path / "cams_krt" / ("frame%05d_cam.txt" % k)), "w") as f:
            for M in (c.K, c.R):
                f.write("\n".join(" ".join("%.9g" % v for v in row) for row in M) + "\n\n")
            f.write(" ".join("%.9g" % v for v in c.t.ravel()) + "\n")
    with open(str(tmp_path / "scene_info.xml"), "w") as f:
        f.write('<bwm_info_for_boxm2><bbox maxx="1.0" maxy="1" maxz="1.5" minx="-1" miny="-1.0 " minz="-0.5"></bbox>'
                '<resolution val="0.01"></resolution></bwm_info_for_boxm2>')
    s = RestrepoScene(str(tmp_path))
    assert s.n_images == n and s.image_shape == (H, W)
    assert np.array_equal(s.bbox, np.array([[-1, -1, -0.5, 1, 1, 1.5]], np.float32))
    im = s.get_image(2)
    assert im.image.dtype == np.float32 and np.array_equal(im.image, pix[2].astype(np.float32) / np.float32(255.))
    assert np.allclose(im.camera.K, cams[2].K, rtol=1e-6) and np.allclose(im.camera.t, cams[2].t, rtol=1e-6, atol=1e-7)
    assert [int(j) for j in s._get_neighbor_idxs(0, 4)] == [1, 2, 3, 4]
    assert len(s.get_image_with_neighbors(3)) == 5 and s.get_image_with_neighbors(3)[0] is s.get_image(3)
    assert s.voxel_grid(np.array([4, 4, 4])).shape == (3, 4, 4, 4)
    sd = RestrepoScene(str(tmp_path), select_neighbors_based_on="distance")
    assert len(set(int(j) for j in sd._get_neighbor_idxs(0, 3)) - {0}) == 3
    ref_scene = "/root/reference/tests/restrepo_mock_dataset/scene_1"
    if os.path.isdir(ref_scene):
        g = np.load(GOLDEN_SCENE)
        r = RestrepoScene(ref_scene)
        assert np.array_equal(parse_scene_info(os.path.join(ref_scene, "scene_info.xml")), g["bbox"])
        assert r.n_images == g["K"].shape[0] and r.image_shape == (720, 1280)
        for i in range(r.n_images):
            c = r.get_image(i).camera
            assert np.array_equal(c.K, g["K"][i]) and np.array_equal(c.R, g["R"][i]) and np.array_equal(c.t, g["t"][i])
            assert np.array_equal(c.P, g["P"][i]) and np.array_equal(c.center, g["center"][i])


def test_dtu_scene_on_disk(tmp_path):
    """A scene written in the DTU MVS layout (tests/golden/make_dtu_golden.py holds the writer and the inputs)
    reads back to what the reference's own DTUScene methods produced on it: K / R / t per view, the bounding
    box from the ObsMask file, ground-truth distance maps, per-pixel depths; view 50 and the other illumination
    settings are ignored."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_dtu_golden import write_scene
    from raynet_b200.common.scene import DTUScene
    g = np.load(os.path.join(ROOT, "tests", "golden", "dtu_golden.npz"))
    write_scene(str(tmp_path), 7, g["K_in"], list(g["P_in"]), g["pix"], g["depths"], g["bb"])
    s = DTUScene(str(tmp_path), 7)
    n = g["P_in"].shape[0]
    assert s.n_images == n and s.image_shape == tuple(g["pix"].shape[1:3])
    assert np.array_equal(s.bbox, g["bbox"]) and s.bbox.dtype == np.float32
    assert s.observation_mask.shape == (2, 2, 2)
    for i in range(n):
        im = s.get_image(i)
        assert np.array_equal(im.camera.K, g["K"][i])
        assert np.allclose(im.camera.R, g["R"][i], rtol=0, atol=1e-6) and np.allclose(im.camera.t, g["t"][i], rtol=1e-6)
        assert np.allclose(im.camera.center, g["center"][i], rtol=1e-5, atol=1e-3)
        assert np.array_equal(im.image, g["pix"][i].astype(np.float32) / np.float32(255.))
        assert np.allclose(s.get_depth_map(i), g["depth_maps"][i], rtol=1e-5, atol=1e-3)
    for (i, y, x), want in zip(g["pixel_queries"], g["pixel_depths"]):
        got = s.get_depth_for_pixel(int(i), int(y), int(x))
        assert (got is None and np.isnan(want)) or abs(got - want) <= 1e-3
    assert s.view_order(1, neighbors=2) == [1, 0, 2]
    assert len(s.get_image_with_neighbors(1, neighbors=2)) == 3
    with pytest.raises(NotImplementedError):
        s.get_pointcloud()
