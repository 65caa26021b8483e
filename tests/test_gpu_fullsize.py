"""GPU tests at BASELINE.json's full sizes (C2: 128^3 / 5 views / 32 planes / 256^2 px / 3 sweeps and the
headline C3: 256^3 / 9 views / 64 planes / 512^2 px / 5 sweeps), where the CPU oracle cannot run the
whole workload in seconds.  They check

  * a random SAMPLE of rays of the full-size run against the oracle, ray by ray: voxel lists
    bit-exact, S_voxel_space within 1e-5, and -- because a ray's new message depends only on
    acc_prev, its own s and its own old message -- one sweep and the depth distribution of the
    sampled rays from the engine's own full-size state within 1e-5;
  * size-independent properties of the whole run: conservation (every voxel of acc_new equals the prior plus the float64 sum of
    the messages the sweep wrote for it -- no lost or doubled scatter-add), normalisation of every
    s_hat row, run-to-run agreement (float atomics reorder sums only), depth maps inside the
    camera's [near, far] range of the bounding box.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TOL_P = 1e-5
GAMMA = 0.05
PRIOR = float(np.float32(np.log(GAMMA) - np.log(1 - GAMMA)))


def sigmoid(x):
    x = np.asarray(x, np.float64)
    return 1.0 / (1.0 + np.exp(-x))


class FullRun(object):
    """The bench workload of one rank (world size 1) on the resident engine, kept for inspection."""

    def __init__(self, torch, name):
        import bench
        from raynet_b200.engine import RayPotentialEngine
        from raynet_b200.synth import camera_arrays
        cfg = bench.CONFIGS[name]
        self.cfg = cfg
        H, W, G, V, D, M, I = (cfg[k] for k in ("H", "W", "G", "V", "D", "M", "I"))
        self.scene = bench.make_scene(cfg, 1)
        self.images = list(range(self.scene.n_images))
        self.views = sorted(set(v for i in self.images for v in self.scene.view_order(i)))
        self.features = np.stack([bench.view_features(v, H, W).numpy() for v in self.views])
        dev = torch.device("cuda")
        self.eng = RayPotentialEngine(M, D, V, bench.F, H, W, bench.PADDING, self.scene.bbox.ravel(), (G, G, G),
                                      gamma=GAMMA, max_rays=len(self.images) * H * W, use_distributed=False)
        self.eng.set_voxel_grid(self.scene.voxel_grid())
        self.feats_dev = torch.from_numpy(self.features).to(dev)
        slot = dict((v, k) for k, v in enumerate(self.views))
        ids = torch.arange(H * W, dtype=torch.int32, device=dev)
        self.cams = []
        for i in self.images:
            order = self.scene.view_order(i)
            P, P_inv, centre = camera_arrays([self.scene.get_image(j) for j in order])
            self.cams.append((order, P, P_inv, centre))
            self.eng.add_image(ids, self.feats_dev, torch.from_numpy(P).to(dev), torch.from_numpy(P_inv).to(dev),
                               torch.from_numpy(centre).to(dev),
                               view_ids=torch.tensor([slot[v] for v in order], dtype=torch.int32, device=dev),
                               n_feature_slots=len(self.views))
        self.eng.finalize_frontend()
        self.grid = np.array([G, G, G], np.int32)
        self.bbox = self.scene.bbox.ravel().astype(np.float32)

    def oracle_frontend(self, oracle, image, pixel_ids):
        import bench
        cfg = self.cfg
        order, P, P_inv, centre = self.cams[image]
        vgrid = oracle.voxel_grid(self.bbox, self.grid)
        feats = np.ascontiguousarray(self.features[[self.views.index(v) for v in order]])
        return oracle.frontend(pixel_ids.astype(np.int32), feats, P, P_inv, centre, vgrid, self.bbox, self.grid, cfg["M"],
                               cfg["D"], cfg["V"], bench.F, cfg["H"], cfg["W"], bench.PADDING)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "the gpu-marked tests need a CUDA device"
    return torch


@pytest.fixture(scope="module", params=["c2", "c3"])
def run(request, torch_cuda):
    r = FullRun(torch_cuda, request.param)
    yield r
    del r.eng, r.feats_dev
    torch_cuda.cuda.empty_cache()


def _sample(run, n_per_image, seed):
    rng = np.random.default_rng(seed)
    HW = run.cfg["H"] * run.cfg["W"]
    out = []
    for k in range(len(run.images)):
        px = np.sort(rng.choice(HW, size=n_per_image, replace=False)).astype(np.int64)
        out.append((k, px))
    return out


def test_fullsize_sampled_frontend_vs_oracle(torch_cuda, oracle, run):
    """Voxel lists bit-exact and S_voxel_space / s_hat within 1e-5 on sampled rays of every image."""
    eng, M = run.eng, run.cfg["M"]
    HW = run.cfg["H"] * run.cfg["W"]
    worst = 0.0
    for k, px in _sample(run, 300, seed=11):
        o = run.oracle_frontend(oracle, k, px)
        rows = torch_cuda.from_numpy(k * HW + px).cuda()
        cnt = eng.count[rows].cpu().numpy()
        assert np.array_equal(cnt, o["cnt"])
        for j in range(0, len(px), 100):      # expand the step codes of the sampled rays only
            for q in range(j, min(j + 100, len(px))):
                idx = eng.voxel_indices(start=int(k * HW + px[q]), n=1).cpu().numpy()[0]
                assert np.array_equal(idx, o["idx"][q]), (k, int(px[q]))
        s_hat = eng.s_hat[rows][:, :M].cpu().numpy()
        for q in range(len(px)):
            L = int(cnt[q])
            if L <= 0:
                continue
            ref = np.clip(o["S_vox"][q, :L], 1e-5, 1 - 1e-5)
            ref = ref / ref.sum(dtype=np.float32)
            worst = max(worst, float(np.abs(s_hat[q, :L] - ref).max()))
    print("full-size sampled front end: max |s_hat - oracle| = %.2e" % worst)
    assert worst <= TOL_P


def test_fullsize_sweeps_properties_and_sampled_parity(torch_cuda, oracle, run):
    torch = torch_cuda
    eng, cfg = run.eng, run.cfg
    M, HW = cfg["M"], cfg["H"] * cfg["W"]
    n = eng.n_rays
    valid = (torch.arange(eng.R, device="cuda")[None, :] < eng.count[:n, None])
    # every s_hat row is a distribution
    sums = torch.where(valid, eng.s_hat[:n], torch.zeros((), device="cuda")).sum(dim=1, dtype=torch.float64)
    live = eng.count[:n] > 0
    assert float((sums[live] - 1.0).abs().max()) <= 1e-5

    sample = _sample(run, 200, seed=5)
    fronts = [(k, px, run.oracle_frontend(oracle, k, px)) for k, px in sample]

    for sweep in range(cfg["I"]):
        acc_prev = eng.accumulator().cpu().numpy()                     # row-major copy of the sweep's input
        old = [eng.messages()[torch.from_numpy(k * HW + px).cuda()].cpu().numpy().copy() for k, px, _ in fronts]
        eng.bp_iteration()
        # scatter-add check, voxel by voxel: prior + the float64 sum of every message the sweep wrote
        # for the voxel (a deterministic index_add over the resident offsets) against the float32
        # RED.ADD accumulator -- a lost or doubled update would show up as a difference of order 1
        bp_rows = (eng.count[:n] > 1)[:, None] & valid
        ref = torch.full((eng.GB,), PRIOR, dtype=torch.float64, device="cuda")
        ref.index_add_(0, eng.lin[:n][bp_rows].long(), eng.msgs[:n][bp_rows].double())
        got = eng.acc_prev.double()
        d_acc = float((got - ref).abs().max())
        d_occ = float((torch.sigmoid(got) - torch.sigmoid(ref)).abs().max())
        total = float(eng.msgs[:n][bp_rows].abs().sum(dtype=torch.float64))
        print("sweep %d: max |acc - f64 scatter| = %.2e (probabilities %.2e), sum of |messages| = %.4g"
              % (sweep, d_acc, d_occ, total))
        assert d_acc <= 5e-5 and d_occ <= 2e-6, (sweep, d_acc, d_occ)
        assert abs(float((got - ref).sum())) <= 5e-6 * total
        del ref, got
        # sampled rays: the oracle's sweep from the same input state gives the same new messages
        for (k, px, o), m_old in zip(fronts, old):
            msgs = np.ascontiguousarray(m_old[:, :M]).astype(np.float32)
            scratch = np.full(tuple(run.grid), PRIOR, np.float64)
            oracle.bp_iteration(o["S_vox"], o["idx"], o["cnt"], run.grid, acc_prev.astype(np.float64), scratch, msgs,
                                acc_f64=True)
            got = eng.messages()[torch.from_numpy(k * HW + px).cuda()].cpu().numpy()[:, :M]
            mask = np.arange(M)[None, :] < o["cnt"][:, None]
            err = np.abs(sigmoid(got) - sigmoid(msgs))[mask].max()
            assert err <= TOL_P, (sweep, k, err)

    # depth distribution of the sampled rays from the engine's final state
    acc = eng.accumulator().cpu().numpy()
    S_new = eng.depth_distribution()
    depth = eng.depth().cpu().numpy()
    assert np.isfinite(depth).all()
    for k, px, o in fronts:
        rows = torch.from_numpy(k * HW + px).cuda()
        msgs = eng.messages()[rows].cpu().numpy()[:, :M]
        ref = oracle.depth_distribution(o["S_vox"], o["idx"], o["cnt"], run.grid, acc, msgs, acc_f64=True)
        got = S_new[rows].cpu().numpy()
        assert np.abs(got - ref).max() <= TOL_P
        # depth maps stay between the nearest point and the farthest corner of the bounding box
        centre = run.cams[k][3].ravel()[:3]
        corners = np.array([[run.bbox[i + 3 * ((c >> i) & 1)] for i in range(3)] for c in range(8)], np.float64)
        far = np.linalg.norm(corners - centre[None, :], axis=1).max()
        near = np.linalg.norm(np.clip(centre, run.bbox[:3], run.bbox[3:]) - centre)
        d = depth[k * HW + px]
        hit = o["cnt"] > 0
        assert (d[hit] >= near - 1e-3).all() and (d[hit] <= far + 1e-3).all()


def test_fullsize_rerun_agrees(torch_cuda, run):
    """Float atomics only reorder sums: two complete runs agree to a few ulp of the accumulator."""
    eng, cfg = run.eng, run.cfg
    res = []
    for _ in range(2):
        eng.iterations_done = 0
        eng.acc_prev.fill_(PRIOR)
        eng.run_bp(2)
        res.append(eng.occupancy().clone())
    assert float((res[0] - res[1]).abs().max()) <= 2e-6


def test_c2_whole_workload_end_to_end_vs_oracle(torch_cuda, oracle, parity_log):
    """BASELINE.json configs[1] (C2: 128^3 grid, 5 views, 32 planes, 5 x 256 x 256 = 327 680 rays, 3 sweeps) from
    images' features to occupancy, WHOLE workload against the CPU oracle run the same way (OpenMP, a few seconds):
    every voxel list bit-exact, every voxel of sigma(acc) after all 3 sweeps within 1e-5 of the float64 flavour of
    the reference -- in parity mode and in the fast mode -- and the depth maps equal wherever the arg-max is decided."""
    torch = torch_cuda
    import bench
    from raynet_b200.engine import RayPotentialEngine
    r = FullRun(torch, "c2")
    cfg = r.cfg
    H, W, M, I = cfg["H"], cfg["W"], cfg["M"], cfg["I"]
    HW = H * W
    ids = np.arange(HW, dtype=np.int32)
    fronts = [r.oracle_frontend(oracle, k, ids) for k in range(len(r.images))]
    idx = np.concatenate([f["idx"] for f in fronts])
    cnt = np.concatenate([f["cnt"] for f in fronts])
    S_vox = np.concatenate([f["S_vox"] for f in fronts])
    acc64, msgs64 = oracle.belief_propagation(S_vox, idx, cnt, r.grid, gamma=GAMMA, bp_iterations=I, acc_f64=True)
    occ64 = oracle.occupancy(acc64)
    acc32, _ = oracle.belief_propagation(S_vox, idx, cnt, r.grid, gamma=GAMMA, bp_iterations=I, acc_f64=False)
    flav = float(np.abs(oracle.occupancy(acc32) - occ64).max())
    S_new = oracle.depth_distribution(S_vox, idx, cnt, r.grid, acc64, msgs64, acc_f64=True)
    vgrid = oracle.voxel_grid(r.bbox, r.grid)
    assert np.array_equal(r.eng.count[:r.eng.n_rays].cpu().numpy(), cnt)
    for a in range(0, len(cnt), 65536):
        assert np.array_equal(r.eng.voxel_indices(start=a, n=min(65536, len(cnt) - a)).cpu().numpy(), idx[a:a + 65536])
    rec = {"rays": int(len(cnt)), "reference_f32_vs_f64_flavour": flav}
    dev = torch.device("cuda")
    for parity in (False, True):
        if parity:
            del r.eng
            torch.cuda.empty_cache()
            eng = RayPotentialEngine(M, cfg["D"], cfg["V"], bench.F, H, W, bench.PADDING, r.bbox, tuple(r.grid), gamma=GAMMA,
                                     max_rays=len(cnt), use_distributed=False, parity=True)
            eng.set_voxel_grid(r.scene.voxel_grid())
            slot = dict((v, k) for k, v in enumerate(r.views))
            pix = torch.arange(HW, dtype=torch.int32, device=dev)
            for (order, P, P_inv, centre) in r.cams:
                eng.add_image(pix, r.feats_dev, torch.from_numpy(P).to(dev), torch.from_numpy(P_inv).to(dev),
                              torch.from_numpy(centre).to(dev),
                              view_ids=torch.tensor([slot[v] for v in order], dtype=torch.int32, device=dev),
                              n_feature_slots=len(r.views))
            eng.finalize_frontend()
        else:
            eng = r.eng
        eng.run_bp(I)
        err = float(np.abs(eng.occupancy().cpu().numpy() - occ64).max())
        depth = eng.depth().cpu().numpy()
        agree = []
        for k, f in enumerate(fronts):
            sl = slice(k * HW, (k + 1) * HW)
            ref_depth, _ = oracle.argmax_depth(S_new[sl], f["idx"], vgrid, r.grid, r.cams[k][3])
            top2 = -np.sort(-S_new[sl], axis=1)[:, :2]
            decided = (top2[:, 0] - top2[:, 1]) > 1e-5
            assert np.abs(depth[sl][decided] - ref_depth[decided]).max() < 1e-6
            agree.append(float((np.abs(depth[sl] - ref_depth) < 1e-6).mean()))
        rec["parity" if parity else "fast"] = {"occupancy": err, "depth_maps_equal_fraction": float(np.mean(agree))}
        print("C2 whole workload, %s: max |occ - f64 ref| over %d voxels = %.2e (f32 ref vs f64 ref %.2e), depth equal %.5f"
              % ("parity" if parity else "fast", occ64.size, err, flav, np.mean(agree)))
        assert err <= TOL_P, (parity, err)
        assert np.mean(agree) > 0.999
    parity_log["c2_whole_workload/3_sweeps"] = rec
