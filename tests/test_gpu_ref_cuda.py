"""The sm_100a kernels against the REFERENCE'S OWN CUDA kernels executed on the same GPU.

oracle/build_ref_cuda.py compiles raynet/cuda_implementations/*.cu + the kernel text of raynet_fp.py
(unchanged but for the zero-initialisation of S, SURVEY.md 2.2 defect 1) into cubins; oracle/ref_cuda.py
launches `batch_raynet_fp` / `batch_complete_depth_estimation` the way raynet_fp.py:275-376 does.  This
pins the three functions that have no CPU twin and no test in the reference -- sample_in_bbox
(sampling_schemes.cu:44-90), the plane-sweep similarity + softmax (feature_similarities.cu:66-124) and
arg-max -> depth (raynet_fp.py:193-226) -- against an execution of the reference itself, and cross-checks
the C oracle's restatement of them.

What can differ, by construction (SURVEY.md 7 "hard parts", 2.2 defects 5-6): the reference CUDA evaluates
bbox / bin size from decimal text in double and is compiled with -fmad=true, so a grazed voxel or a pixel
that sits on a rounding boundary can flip (the bit-exact target is the Cython flavour, and one flipped
(plane, view) sample changes that ray's whole distribution); its BP uses `total - inclusive prefix` for
the suffix sums and does not skip count <= 1 rays.  The tests therefore require identical voxel lists for
>= 99 % of the rays and a distribution within 1e-5 for >= 99.5 % of those, and compare the BP / depth
values on the rays that agree.
"""
import numpy as np
import pytest

from rig import case_c1, case_nine, case_small, sigmoid

pytestmark = pytest.mark.gpu

TOL_P = 1e-5
PRIOR = float(np.float32(np.log(0.05) - np.log(1 - 0.05)))


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name,mk", [("c1", case_c1), ("small", case_small), ("nine", case_nine)])
def test_against_reference_cuda_kernels(torch_cuda, oracle, name, mk):
    torch = torch_cuda
    from oracle import ref_cuda
    if not ref_cuda.available(name):
        pytest.skip("oracle/_ref/cuda/raynet_fp_%s.cubin not built (needs /root/reference at build time)" % name)
    from raynet_b200.cuda_implementations.raynet_fp import perform_raynet_fp
    c = mk()
    ref = ref_cuda.RefCuda(name)
    assert ref.params["M"] == c.M and ref.params["D"] == c.D and ref.params["N"] == c.V and ref.params["H"] == c.H
    ins = [_dev(torch, x) for x in (c.ray_idxs, c.features.ravel(), c.P.ravel(), c.P_inv.ravel(), c.centre, c.vgrid.ravel())]

    def buffers():
        return dict(idx=torch.zeros((c.N, c.M, 3), dtype=torch.int32, device="cuda"),
                    cnt=torch.zeros((c.N,), dtype=torch.int32, device="cuda"),
                    S=torch.zeros((c.N, c.M), dtype=torch.float32, device="cuda"),
                    msgs=torch.zeros((c.N, c.M), dtype=torch.float32, device="cuda"),
                    acc_in=torch.full(tuple(c.grid), PRIOR, dtype=torch.float32, device="cuda"),
                    acc_out=torch.full(tuple(c.grid), PRIOR, dtype=torch.float32, device="cuda"))

    # ---- sweep 1 from the prior: reference kernel vs ours through the same closure signature ----
    r, g = buffers(), buffers()
    ref.raynet_fp(*ins, r["idx"], r["cnt"], r["S"], r["acc_in"], r["msgs"], r["acc_out"])
    fp, de = perform_raynet_fp(c.M, c.D, c.V, 32, c.H, c.W, 11, c.bbox, c.grid, "sample_in_bbox")
    fp(*ins, g["idx"], g["cnt"], g["S"], g["acc_in"], g["msgs"], g["acc_out"])
    torch.cuda.synchronize()
    rc, gc = r["cnt"].cpu().numpy(), g["cnt"].cpu().numpy()
    ri, gi = r["idx"].cpu().numpy(), g["idx"].cpu().numpy()
    same = (rc == gc) & np.all(ri == gi, axis=(1, 2))
    print("%s: voxel lists identical for %d of %d rays" % (name, int(same.sum()), c.N))
    assert same.mean() >= 0.99
    # the reference kernel leaves clip_and_renorm(S_voxel_space) in S (mrf_bp.cu:103-112); ours leaves S_voxel_space
    gS = g["S"].cpu().numpy()
    rS = r["S"].cpu().numpy()
    valid = np.arange(c.M)[None, :] < gc[:, None]
    clipped = np.where(valid, np.clip(gS, 1e-5, 1 - 1e-5), 0).astype(np.float32)
    g_hat = clipped / np.maximum(clipped.sum(axis=1, keepdims=True, dtype=np.float32), 1e-30)
    live = same & (gc > 0)
    ray_err = np.where(valid, np.abs(g_hat - rS), 0).max(axis=1)
    agree = live & (ray_err <= TOL_P)       # rays on which no (plane, view) sample flipped its rounded pixel
    print("%s: clip_and_renorm(S_voxel_space) within 1e-5 of the reference for %d of %d rays (max on those %.2e)"
          % (name, int(agree.sum()), int(live.sum()), float(ray_err[agree].max())))
    assert agree.sum() >= 0.995 * live.sum()
    # first-sweep messages and accumulator (probabilities), rays BP touches in both implementations
    bp = agree & (gc > 1)
    gm, rm = g["msgs"].cpu().numpy(), r["msgs"].cpu().numpy()
    err_m = float(np.abs(sigmoid(gm[bp]) - sigmoid(rm[bp]))[valid[bp]].max())
    print("%s: max |sigma(msg) - reference| = %.2e" % (name, err_m))
    assert err_m <= TOL_P
    if bool(same.all()) and bool((gc != 1).all()) and bool(agree.sum() == live.sum()):      # identical inputs feed the accumulator
        err_a = float(np.abs(sigmoid(g["acc_out"].cpu().numpy()) - sigmoid(r["acc_out"].cpu().numpy())).max())
        print("%s: max |sigma(acc) - reference| = %.2e" % (name, err_a))
        assert err_a <= TOL_P

    # ---- depth estimation from a common state (ours after the sweep) ---------------------------
    acc = g["acc_out"].clone()
    msgs = g["msgs"].clone()
    r2, g2 = buffers(), buffers()
    d_ref = torch.zeros((c.N,), dtype=torch.float32, device="cuda")
    d_got = torch.zeros((c.N,), dtype=torch.float32, device="cuda")
    ref.raynet_de(*ins, r2["idx"], r2["cnt"], r2["S"], acc, msgs, d_ref)
    de(*ins, g2["idx"], g2["cnt"], g2["S"], acc, msgs, d_got)
    torch.cuda.synchronize()
    rS2, gS2 = r2["S"].cpu().numpy(), g2["S"].cpu().numpy()        # S_new, the re-estimated depth distribution
    err_d = float(np.abs(gS2[bp] - rS2[bp]).max())
    print("%s: max |S_new - reference| = %.2e" % (name, err_d))
    assert err_d <= TOL_P
    top2 = -np.sort(-rS2, axis=1)[:, :2]
    decided = bp & ((top2[:, 0] - top2[:, 1]) > 1e-5)
    dr, dg = d_ref.cpu().numpy(), d_got.cpu().numpy()
    assert decided.sum() > 0.8 * bp.sum()
    assert np.abs(dr[decided] - dg[decided]).max() < 1e-5


def test_reference_cuda_speed_bar(torch_cuda):
    """The "reference kernel on the same box" bar of SURVEY.md 8d: one reference image of the headline
    configuration C3 through the reference's own `batch_raynet_fp` (front end + one BP sweep: what the
    reference launches per image per sweep, forward_pass.py:650-663) and `batch_complete_depth_estimation`
    (:723-736), timed with CUDA events; a complete inference of the reference is I x images fp launches +
    images de launches.  Prints the figure quoted in profiles/README.md and checks that the resident
    pipeline is at least an order of magnitude faster on the same workload."""
    import json
    import os
    torch = torch_cuda
    import bench
    from oracle import ref_cuda
    from raynet_b200.engine import RayPotentialEngine
    from raynet_b200.synth import camera_arrays
    if not ref_cuda.available("c3"):
        pytest.skip("oracle/_ref/cuda/raynet_fp_c3.cubin not built")
    cfg = bench.CONFIGS["c3"]
    H, W, G, V, D, M, I = (cfg[k] for k in ("H", "W", "G", "V", "D", "M", "I"))
    dev = torch.device("cuda")
    scene = bench.make_scene(cfg, 1)
    order = scene.view_order(0)
    P, P_inv, centre = camera_arrays([scene.get_image(j) for j in order])
    feats = torch.stack([bench.view_features(v, H, W) for v in order]).to(dev).contiguous()
    n = H * W
    ids = torch.arange(n, dtype=torch.int32, device=dev)
    dP, dPi, dC = torch.from_numpy(P).to(dev), torch.from_numpy(P_inv).to(dev), torch.from_numpy(centre).to(dev)
    vgrid = torch.from_numpy(np.ascontiguousarray(scene.voxel_grid().transpose(1, 2, 3, 0))).to(dev)
    ref = ref_cuda.RefCuda("c3")
    ins = [ids, feats.reshape(-1), dP.reshape(-1), dPi.reshape(-1), dC.reshape(-1), vgrid.reshape(-1)]
    idx = torch.zeros((n, M, 3), dtype=torch.int32, device=dev)
    cnt = torch.zeros((n,), dtype=torch.int32, device=dev)
    S = torch.zeros((n, M), dtype=torch.float32, device=dev)
    msgs = torch.zeros((n, M), dtype=torch.float32, device=dev)
    acc_in = torch.full((G, G, G), PRIOR, dtype=torch.float32, device=dev)
    acc_out = torch.full((G, G, G), PRIOR, dtype=torch.float32, device=dev)
    depth = torch.zeros((n,), dtype=torch.float32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ref.raynet_fp(*ins, idx, cnt, S, acc_in, msgs, acc_out)          # warm-up
    torch.cuda.synchronize()
    ev[0].record()
    ref.raynet_fp(*ins, idx, cnt, S, acc_in, msgs, acc_out)
    ev[1].record()
    ref.raynet_de(*ins, idx, cnt, S, acc_out, msgs, depth)
    ev[2].record()
    torch.cuda.synchronize()
    t_fp, t_de = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    ref_rays_per_s = V * n / (V * (I * t_fp + t_de) * 1e-3)
    del idx, S, msgs
    # ours: the same image through the resident pipeline (front end + I sweeps + depth)
    eng = RayPotentialEngine(M, D, V, bench.F, H, W, bench.PADDING, scene.bbox.ravel(), (G, G, G), gamma=bench.GAMMA,
                             max_rays=n, use_distributed=False)
    eng.set_voxel_grid(scene.voxel_grid())

    def ours():
        eng.reset()
        eng.add_image(ids, feats, dP, dPi, dC)
        eng.run_bp(I)
        return eng.depth()

    ours()
    torch.cuda.synchronize()
    ev[0].record()
    d_ours = ours()
    ev[1].record()
    torch.cuda.synchronize()
    ours_rays_per_s = n / (ev[0].elapsed_time(ev[1]) * 1e-3)
    out = {"reference_cuda_rays_per_s": ref_rays_per_s, "fp_launch_ms": t_fp, "de_launch_ms": t_de,
           "ours_rays_per_s_one_image": ours_rays_per_s, "ratio": ours_rays_per_s / ref_rays_per_s,
           "what": "C3, one reference image (%d rays); reference = its own CUDA kernels built for sm_100a, %d sweeps x fp + de" % (n, I)}
    print("reference CUDA speed bar: " + json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(os.path.join("gpurun_out", "ref_cuda_speed_bar.json"), "w"), indent=1)
    assert np.isfinite(d_ours.cpu().numpy()).all()
    assert ours_rays_per_s > 10 * ref_rays_per_s
