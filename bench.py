#!/usr/bin/env python
"""bench.py -- rays/s of complete ray-potential inference (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config c3] [--scaling strong|weak] [--impl reference]
    (N > 1: launched by torchrun, one rank per GPU, NCCL)

One "step" = one complete inference over the job:
    front end (sample_in_bbox + plane-sweep similarity + DDA + plane->voxel) once per ray,
    I synchronous BP sweeps over all rays (+ one all-reduce of the occupancy accumulator per
    sweep when N > 1), one depth pass (+ arg-max -> depth).
Scaling.  BASELINE.json configs[3] is "the 256^3 workload sharded by ray batch across 2/4/8 GPUs":
    STRONG scaling (default for c1/c2/c3) -- the job is fixed (C3: 9 reference images x 512 x 512 rays); the
    (image, column-major pixel) ray enumeration is cut into N contiguous blocks (raynet_b200/sharding.py),
    so a rank owns whole and partial images.  --scaling weak (default for c5, whose 16 images only fit on
    8 GPUs): every rank owns the same number of whole reference images.
`value`  : rays/s, whole job, inputs (feature maps, cameras, ray ids) resident in HBM.
`e2e`    : the same through the reference-facing plug-in call
           RayNetForwardPass.forward_pass(scene, images_range) with HOST IMAGES in (zero-padded views from
           pinned memory -> MV-CNN on the device, raynet_b200.models.SimpleCNN) and HOST depth maps out
           (H2D / D2H inside the timed region), timed over all --steps.
`roofline`: the BP sweep (bp4_kernel), first and non-first sweeps separately (a first sweep moves 12 B per
           traversed voxel, the others 20 B: SURVEY.md 8d), plus per-stage entries for the front end and
           the depth pass against the same byte model.

--impl reference : the CPU implementation of the path (oracle port; see cpu_baseline.kind) on
the host cores (all of them, also under torchrun), bounded sample per step.

Extra keys of the N = 1 line (reported baselines / neighbours, none of them part of `value`):
`cpu_baseline` (the CPU port on ~15 s of the same workload) and `cnn` (the MV-CNN feature extractor of
SURVEY.md 8(f) row 1 on this rank's views).  The speed of the reference's OWN CUDA kernels on the same GPU
is measured by tests/test_gpu_ref_cuda.py::test_reference_cuda_speed_bar (only tests may run oracle/_ref/cuda).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: grid, views, planes, H, W, M, sweeps   
    "c1": dict(G=32, V=2, D=16, H=64, W=64, M=96, I=3, label="C1 32^3 grid, 2 views, 16 planes, 64x64 px, 3 sweeps"),
    "c2": dict(G=128, V=5, D=32, H=256, W=256, M=384, I=3,
               label="C2 DTU-style 128^3 grid, 5 views, 32 planes, 256x256 px, 3 sweeps"),
    "c3": dict(G=256, V=9, D=64, H=512, W=512, M=768, I=5,
               label="C3 headline 256^3 grid, 9 views, 64 planes, 512x512 px, 5 sweeps"),
    # BASELINE.json configs[4], meant for 8 GPUs: a ring of 16 cameras, every image with its 14 next neighbours,
    # two reference images per GPU (the per-ray state of one 1024^2 image at M = 1536 is 19 GB)
    "c5": dict(G=512, V=15, D=128, H=1024, W=1024, M=1536, I=5, images_per_gpu=2, ring=16,
               label="C5 aerial-style 512^3 grid, 15 views, 128 planes, 1024x1024 px, 5 sweeps, 2 images per GPU"),
}
F, PADDING, GAMMA = 32, 11, 0.05
METRIC = "rays/s, complete ray-potential inference (front end + I BP sweeps + depth)"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic():
    """DRAM bytes of one non-first BP sweep from the committed ncu capture of this configuration
    (profiles/roofline_traffic.json).  NOT measured by this run -- ncu cannot run inside the timed
    region; the line says so next to the number."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------
def view_features(view, H, W, pinned=False):
    """f32 [H+p+1, W+p+1, F] N(0,1)/sqrt(F), border row/col 0 zeroed; seeded by the view id."""
    import torch
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + int(view))
    f = torch.randn((H + PADDING + 1, W + PADDING + 1, F), generator=g, dtype=torch.float32)
    f /= float(np.sqrt(F))
    f[0, :, :] = 0
    f[:, 0, :] = 0
    return f


class FeatureModel(object):
    """Stands for the MV-CNN (out of scope): hands out precomputed per-view feature maps that
    live in pinned host memory, like a feature cache would."""

    def __init__(self, views, H, W):
        import torch
        self.views = list(views)
        self.host = torch.empty((len(self.views), H + PADDING + 1, W + PADDING + 1, F), dtype=torch.float32)
        for k, v in enumerate(self.views):
            self.host[k].copy_(view_features(v, H, W))
        if torch.cuda.is_available():
            self.host = self.host.pin_memory()

    def predict_features(self, scene, view_indices):
        if list(view_indices) == self.views:
            return self.host
        import torch
        return self.host[torch.tensor([self.views.index(v) for v in view_indices])]


def scaling_of(args, cfg):
    if args.scaling:
        return args.scaling
    return "weak" if "images_per_gpu" in cfg else "strong"


def make_scene(cfg, world, scaling="weak", with_images=False):
    """strong: the fixed job (a ring of V cameras, every view a reference image with the others as neighbours;
    c5: its ring of 16).  weak: V reference images per GPU on a ring of V * world cameras (rank r owns images
    r, r + world, ... whose neighbours are the same residue class)."""
    from raynet_b200.synth import SyntheticScene
    if "ring" in cfg:      # fixed ring: neighbours are the next V - 1 cameras, images dealt out round-robin
        return SyntheticScene(cfg["ring"], cfg["H"], cfg["W"], (cfg["G"],) * 3, neighbors=cfg["V"] - 1, neighbor_stride=1,
                              with_images=with_images)
    if scaling == "strong":
        return SyntheticScene(cfg["V"], cfg["H"], cfg["W"], (cfg["G"],) * 3, neighbors=cfg["V"] - 1, neighbor_stride=1,
                              with_images=with_images)
    n_total = cfg["V"] * world
    return SyntheticScene(n_total, cfg["H"], cfg["W"], (cfg["G"],) * 3, neighbors=cfg["V"] - 1,
                          neighbor_stride=world, with_images=with_images)


def segments_of_rank(cfg, scene, rank, world, scaling):
    """This rank's share of the job as [(image, first ray, last ray)]."""
    from raynet_b200 import sharding
    H, W = cfg["H"], cfg["W"]
    if scaling == "strong":
        n = scene.n_images
        unit = 64 * H if (H * W) % (64 * H) == 0 else (8 * H if (H * W) % (8 * H) == 0 else 1)
        return sharding.image_segments([H * W] * n, rank, world, unit)
    imgs = list(range(rank, scene.n_images, world))
    if "images_per_gpu" in cfg:
        imgs = imgs[:cfg["images_per_gpu"]]
    return [(i, 0, H * W) for i in imgs]


def start_clock_sampler(device_index):
    import torch
    uuid = "GPU-" + str(torch.cuda.get_device_properties(device_index).uuid)
    out = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        proc = subprocess.Popen(["nvidia-smi", "-i", uuid, "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                 "-lms", "50"], stdout=out, stderr=subprocess.DEVNULL)
    except Exception:
        return None, out.name
    return proc, out.name


def stop_clock_sampler(proc, path):
    clocks = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    if proc is not None:
        proc.terminate()
        try:
            proc.wait(timeout=5)
        except Exception:
            proc.kill()
    try:
        rows = [l.strip().split(", ") for l in open(path) if l.strip()]
        sm = [float(r[0]) for r in rows if len(r) >= 8]
        if sm:
            clocks["sm_mhz"] = float(np.median(sm))
            clocks["sm_max_mhz"] = float(rows[0][1])
            clocks["power_w_max"] = max(float(r[2]) for r in rows if len(r) >= 8)
            clocks["samples"] = len(sm)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for k, nm in enumerate(names):
                if any(r[4 + k].strip().lower() == "active" for r in rows if len(r) >= 8):
                    clocks["reasons"].append(nm)
    except Exception as e:
        clocks["error"] = repr(e)
    try:
        os.unlink(path)
    except OSError:
        pass
    return clocks


# ------------------------------------------------------------------------------------------
# CPU arm (oracle port): bounded sample of the same workload
# ------------------------------------------------------------------------------------------
class CpuSample(object):
    def __init__(self, cfg, n_rays):
        from oracle import oracle as orc
        from raynet_b200.synth import camera_arrays
        self.orc = orc
        self.cfg = cfg
        scene = make_scene(cfg, 1, "strong")
        H, W = cfg["H"], cfg["W"]
        order = scene.view_order(0)
        self.P, self.P_inv, self.centre = camera_arrays([scene.get_image(j) for j in order])
        self.features = np.stack([view_features(v, H, W).numpy() for v in order])
        stride = max(1, (H * W) // n_rays)
        self.ray_idxs = np.arange(0, H * W, stride, dtype=np.int32)[:n_rays]     # pixels spread over the image
        self.grid = np.array([cfg["G"]] * 3, np.int32)
        self.bbox = np.array([-1, -1, -1, 1, 1, 1], np.float32)
        self.vgrid = orc.voxel_grid(self.bbox, self.grid)

    def step(self):
        """front end + I sweeps + depth distribution + arg-max depth, all host threads (OpenMP)."""
        c, orc = self.cfg, self.orc
        t0 = time.perf_counter()
        o = orc.frontend(self.ray_idxs, self.features, self.P, self.P_inv, self.centre, self.vgrid, self.bbox,
                         self.grid, c["M"], c["D"], c["V"], F, c["H"], c["W"], PADDING, want_stages=False)
        acc, msgs = orc.belief_propagation(o["S_vox"], o["idx"], o["cnt"], self.grid, gamma=GAMMA,
                                           bp_iterations=c["I"])
        S_new = orc.depth_distribution(o["S_vox"], o["idx"], o["cnt"], self.grid, acc, msgs)
        orc.argmax_depth(S_new, o["idx"], self.vgrid, self.grid, self.centre)
        return time.perf_counter() - t0


def cpu_calibrated_sample(cfg, target_s):
    """Pick a ray count whose step takes about target_s on this host (at most one whole image:
    longer samples repeat the step, see cpu_timed_passes)."""
    n, t = 4096, None
    for _ in range(3):
        probe = CpuSample(cfg, n)
        probe.step()
        t = probe.step()
        if t >= 0.25 * target_s or n >= cfg["H"] * cfg["W"]:
            break
        n = int(min(cfg["H"] * cfg["W"], max(2 * n, n * 0.6 * target_s / max(t, 1e-4))))
    return probe


def cpu_timed_passes(sample, target_s):
    """Repeat sample.step() until about target_s of CPU work has been timed.  Returns (rays, seconds, passes)."""
    rays, secs, passes = 0, 0.0, 0
    while secs < target_s and passes < 64:
        secs += sample.step()
        rays += int(sample.ray_idxs.shape[0])
        passes += 1
    return rays, secs, passes


def cnn_bar(cfg, dev, n_views):
    """the MV-CNN (raynet_b200.models.SimpleCNN, 5 x conv3x3(32) + BN, fp32 CUDA cores)
    on this rank's views, zero-padded images resident on the device; not part of `value` (the metric takes
    feature volumes as inputs)."""
    import torch
    from raynet_b200.models import SimpleCNN
    H, W = cfg["H"], cfg["W"]
    model = SimpleCNN.random_init(channels=3, seed=0)
    g = torch.Generator(device="cpu")
    g.manual_seed(7)
    X = torch.zeros((n_views, H + 2 * PADDING, W + 2 * PADDING, 3), dtype=torch.float32)
    X[:, PADDING:PADDING + H, PADDING:PADDING + W, :] = torch.rand((n_views, H, W, 3), generator=g)
    X = X.to(dev)
    for _ in range(2):
        model.predict_device(X)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 5
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        out = model.predict_device(X)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    flops, h, w, cin = 0.0, H + 2 * PADDING, W + 2 * PADDING, 3
    for _ in range(5):
        h, w = h - 2, w - 2
        flops += 2.0 * n_views * h * w * 9 * cin * 32
        cin = 32
    sm_clock = torch.cuda.get_device_properties(dev).clock_rate * 1e3 if hasattr(torch.cuda.get_device_properties(dev), "clock_rate") else 1.965e9
    peak = torch.cuda.get_device_properties(dev).multi_processor_count * 128 * 2 * sm_clock / 1e12
    return {"ms": ms, "views": n_views, "output_shape": list(out.shape), "tflops": flops / (ms * 1e-3) / 1e12,
            "fp32_peak_tflops_nominal": peak, "frac": flops / (ms * 1e-3) / 1e12 / peak, "launches": 5,
            "what": "5 x (conv3x3 -> 32 channels + folded batch norm [+ ReLU]) on %d zero-padded %dx%dx3 views, fp32 FMA "
                    "pipe; peak = SMs x 128 FMA x 2 x max SM clock" % (n_views, H + 2 * PADDING, W + 2 * PADDING)}


def run_reference_arm(args, cfg):
    """--impl reference: the CPU implementation on the host cores (rank 0 only; the other ranks exit).
    torchrun exports OMP_NUM_THREADS=1 to its workers: the thread count is set explicitly to every core
    this process may run on, so the arm is the same at every N."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    orc.set_threads(host_cores())
    cores = orc.num_threads()
    sample = cpu_calibrated_sample(cfg, 3.0)
    for _ in range(args.warmup):
        sample.step()
    times = [sample.step() for _ in range(args.steps)]
    ms = 1e3 * float(np.mean(times))
    n = int(sample.ray_idxs.shape[0])
    value = n / (ms * 1e-3)
    desc = "%d rays of reference image 0 of %s per step (strided pixels), full pipeline" % (n, cfg["label"])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": scaling_of(args, cfg), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["label"], "rays_per_step": n, "host": "CPU only, %d threads" % cores},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port", "sample": desc,
                         "note": "OpenMP C restatement of the reference's numpy/Cython/.cu path (oracle/rn_oracle.c); "
                                 "the reference's own CPU code is single-threaded Python loops"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_e2e(args, cfg, rank, world, dev, barrier, total_rays, scaling):
    """e2e arm: the reference-facing plug-in call with HOST buffers: images in (pinned zero-padded views ->
    MV-CNN on the device), host depth maps out; every --steps step timed."""
    import torch
    import torch.distributed as dist
    from raynet_b200.common.generation_parameters import GenerationParameters
    from raynet_b200.forward_pass import RayNetForwardPass
    from raynet_b200.models import SimpleCNN
    H, W, G, V, D, M, I = (cfg[k] for k in ("H", "W", "G", "V", "D", "M", "I"))
    scene = make_scene(cfg, world, scaling, with_images=True)
    n_total = scene.n_images
    model = SimpleCNN.random_init(channels=3, seed=0)
    gp = GenerationParameters(depth_planes=D, neighbors=V - 1, grid_shape=np.array([G, G, G], np.int32),
                              max_number_of_marched_voxels=M, padding=PADDING, gamma_mrf=GAMMA)
    if scaling == "strong":
        fp = RayNetForwardPass(model, gp, "sample_in_bbox", scene.image_shape, rays_batch=H * W, bp_iterations=I,
                               shard="rays")
        images_range = (0, n_total, 1)
        n_maps = n_total
    else:
        fp = RayNetForwardPass(model, gp, "sample_in_bbox", scene.image_shape, rays_batch=H * W, bp_iterations=I,
                               shard="images")
        stop = n_total if "images_per_gpu" not in cfg else min(n_total, rank + world * cfg["images_per_gpu"])
        images_range = (rank, stop, world)
        n_maps = len(range(rank, stop, world))

    def e2e_step():
        maps = list(fp.forward_pass(scene, images_range))
        assert len(maps) == n_maps and maps[0].shape == (H, W)
        return maps

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())
    return {"value": total_rays / (e2e_ms * 1e-3), "unit": "rays/s", "ms_per_step": e2e_ms, "steps_timed": args.steps,
            "h2d_bytes_per_step": int(fp.h2d_bytes), "d2h_bytes_per_step": int(fp.d2h_bytes),
            "api": "raynet_b200.forward_pass.RayNetForwardPass.forward_pass(scene, images_range): zero-padded host "
                   "images in (pinned) -> SimpleCNN on the device -> ray-potential inference -> host depth maps "
                   "out; host wall clock around the call, barrier + synchronize on both sides, max over ranks"}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_gpu_arm(args, cfg):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this path has no CPU fallback (use --impl reference for "
                         "the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print("[bench] note: --gpus %d but WORLD_SIZE=%d; using WORLD_SIZE" % (args.gpus, world), file=sys.stderr)

    from raynet_b200 import _lib
    from raynet_b200.engine import RayPotentialEngine
    from raynet_b200.synth import camera_arrays
    _lib.load()

    scaling = scaling_of(args, cfg)
    H, W, G, V, D, M, I = (cfg[k] for k in ("H", "W", "G", "V", "D", "M", "I"))
    scene = make_scene(cfg, world, scaling)
    dev = torch.device("cuda", local)
    ids = torch.arange(H * W, dtype=torch.int32, device=dev)

    def setup(segs, capacity):
        """Device-resident inputs of this rank's segments [(image, first ray, last ray)] + the engine."""
        my_images = sorted(set(i for (i, _, _) in segs))
        my_views = sorted(set(v for i in my_images for v in scene.view_order(i)))
        model = FeatureModel(my_views, H, W)
        eng = RayPotentialEngine(M, D, V, F, H, W, PADDING, scene.bbox.ravel(), (G, G, G), gamma=GAMMA, max_rays=capacity,
                                 collective=args.collective, fuse_first_sweep=not args.no_fuse)
        eng.set_voxel_grid(scene.voxel_grid())
        feats = model.host.to(dev)
        slot = dict((v, k) for k, v in enumerate(my_views))
        per_seg = []
        for (i, a, b) in segs:
            order = scene.view_order(i)
            P, P_inv, centre = camera_arrays([scene.get_image(j) for j in order])
            per_seg.append((ids[a:b], torch.from_numpy(P).to(dev), torch.from_numpy(P_inv).to(dev),
                            torch.from_numpy(centre).to(dev),
                            torch.tensor([slot[v] for v in order], dtype=torch.int32, device=dev)))
        return eng, feats, per_seg, my_views

    segs = segments_of_rank(cfg, scene, rank, world, scaling)       # this rank's (image, first ray, last ray)
    n_rays = int(sum(b - a for (_, a, b) in segs))
    eng, feats, per_seg, my_views = setup(segs, n_rays)
    balanced = False
    if world > 1 and scaling == "strong" and (H * W) % (8 * H) == 0 and not args.no_balance:
        # blocks of equal WORK instead of equal ray counts (rays near the image border cross few voxels): one
        # tracing pass gives the traversed voxels per group of 8 image columns, summed over the ranks; the plan is
        # what RayNetForwardPass caches per job after its first call (raynet_b200/sharding.py)
        from raynet_b200 import sharding
        for (sid, P, P_inv, centre, vids) in per_seg:
            eng.trace_image(sid, P_inv, centre)
        unit = 8 * H
        lens = [H * W] * scene.n_images
        work = torch.zeros((sum(lens) // unit,), dtype=torch.float64, device=dev)
        lo_u = (segs[0][0] * H * W + segs[0][1]) // unit
        work[lo_u:lo_u + n_rays // unit] = eng.unit_work(unit)
        dist.all_reduce(work)
        bounds = sharding.balanced_boundaries(work.cpu().numpy(), world)
        all_segs = [sharding.segments_from_unit_boundaries(lens, unit, bounds[r], bounds[r + 1]) for r in range(world)]
        segs = all_segs[rank]
        n_rays = int(sum(b - a for (_, a, b) in segs))
        del eng, feats, per_seg
        torch.cuda.empty_cache()
        eng, feats, per_seg, my_views = setup(segs, max(sum(b - a for (_, a, b) in sg) for sg in all_segs))
        balanced = True
    stage_events = []

    def device_step(record=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        if record:
            ev[0].record()
        eng.reset()
        # same order as RayNetForwardPass.forward_pass: trace every segment, bin the rays (the class sizes
        # travel to the host on a side stream meanwhile), then similarity + mapping per segment
        for (sid, P, P_inv, centre, vids) in per_seg:
            eng.trace_image(sid, P_inv, centre)
        eng.finalize_frontend()
        for k, (sid, P, P_inv, centre, vids) in enumerate(per_seg):
            eng.score_image(k, feats, P, view_ids=vids, n_feature_slots=len(my_views))
        if record:
            ev[1].record()
        eng.run_bp(I)
        if record:
            ev[2].record()
        depth = eng.depth()
        if record:
            ev[3].record()
            stage_events.append(ev)
        return depth

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            fn()
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    for _ in range(max(args.warmup, 3)):
        device_step()
    proc, path = start_clock_sampler(local) if rank == 0 else (None, None)
    eng.sweep_events = []
    eng.exchange_events = [] if world > 1 else None
    launches0 = eng.launches
    ms_step = timed(lambda: device_step(record=True), args.steps)
    gpu_launches = (eng.launches - launches0) // args.steps
    sweep_ms = np.array([a.elapsed_time(b) for (a, b) in eng.sweep_events]).reshape(args.steps, I)
    eng.sweep_events = None
    exchange_ms = None
    if eng.exchange_events:
        exchange_ms = float(np.mean([a.elapsed_time(b) for (a, b) in eng.exchange_events]))
    eng.exchange_events = None
    clocks = stop_clock_sampler(proc, path) if rank == 0 else None
    stages = np.array([[e[k].elapsed_time(e[k + 1]) for k in range(3)] for e in stage_events])
    counts = eng.count[:eng.n_rays]
    sum_L = int(counts[counts > 1].sum().item())
    mean_L = float(counts.float().mean().item()) if eng.n_rays else 0.0
    max_L = int(eng.max_count)
    totals = torch.tensor([n_rays], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(totals)
    total_rays = int(totals.item())
    value = total_rays / (ms_step * 1e-3)

    collective = eng.collective
    e2e = None
    del eng, feats
    torch.cuda.empty_cache()
    if not args.no_e2e:
        e2e = run_e2e(args, cfg, rank, world, dev, barrier, total_rays, scaling)
    cnn = None
    if rank == 0 and not args.no_cpu:
        torch.cuda.empty_cache()
        try:
            cnn = cnn_bar(cfg, dev, len(my_views))
        except Exception as e:
            cnn = {"unavailable": repr(e)}

    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel: the BP sweep.  Algorithmic bytes per traversed voxel: a non-first sweep
        # moves s_hat 4 + msg in 4 + msg out 4 + acc gather 4 + acc scatter-add 4 = 20 B; the first sweep after
        # a reset reads no messages and gathers nothing (the accumulator is the prior everywhere) = 12 B
        first_ms = float(sweep_ms[:, 0].mean())
        next_ms = float(sweep_ms[:, 1:].mean()) if I > 1 else first_ms
        b_first, b_next = 12.0 * sum_L, 20.0 * sum_L
        a_next = b_next / (next_ms * 1e-3) / 1e9
        a_first = b_first / (first_ms * 1e-3) / 1e9
        blend_bytes = b_first + (I - 1) * b_next
        blend_ms = float(sweep_ms.sum(axis=1).mean())
        traffic = profiled_traffic() if args.config == "c3" and world == 1 else None
        # per-stage byte model of SURVEY.md 8d (per GPU): front end = 24 B/ray + 4 B per voxel (S_vox written) + the
        # feature maps once; depth = 12 B per voxel + 4 B per ray; grid work = 3 x 4 B x G^3 per sweep
        feat_bytes = len(my_views) * (H + PADDING + 1) * (W + PADDING + 1) * F * 4.0
        # (fused first sweep: the 4 B/voxel S_vox write of the front end happens inside the first sweep, which in turn
        # no longer reads s_hat -- its 12 B/voxel stay 12 -- and the front end writes 4 D bytes of plane scores per ray)
        fe_bytes = 24.0 * n_rays + feat_bytes + ((4.0 * D * n_rays) if not args.no_fuse else 4.0 * sum_L)
        de_bytes = 12.0 * sum_L + 4.0 * n_rays
        grid_bytes = I * 3 * G ** 3 * 4.0
        fe_ms, bp_ms, de_ms = (float(stages[:, k].mean()) for k in range(3))
        step_bytes = fe_bytes + blend_bytes + de_bytes + grid_bytes

        def frac(nbytes, ms):
            return nbytes / (ms * 1e-3) / 1e9 / peak

        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": cfg["label"], "rays_total": total_rays, "rays_on_rank0": n_rays,
                "segments_on_rank0": [[int(i), int(a), int(b)] for (i, a, b) in segs], "bp_sweeps": I, "max_voxels": M,
                "mean_voxels_per_ray": mean_L, "longest_ray": max_L,
                "parallelism": ("rays sharded in %d contiguous blocks of the (image, pixel) enumeration, %s" %
                                (world, "cut for equal traversed voxels" if balanced else "equal ray counts"))
                               if scaling == "strong" else "whole reference images per rank, dp%d" % world,
                "first_sweep": "plane->voxel mapping fused into the first sweep (rn_engine_first_sweep_mapped)"
                               if not args.no_fuse else "separate mapping kernel",
                "collective": ("sum of the per-rank partial f32[%d^3] accumulators after every sweep: %s" % (G, {
                    "peer": "this library's fused barrier + reduce + broadcast kernel over NVLink peer memory "
                            "(rn_peer_allreduce_f32)",
                    "peer_multicast": "this library's fused exchange kernel, sum and broadcast inside the NVSwitch "
                                      "(multimem.ld_reduce / multimem.st on the multicast addresses, "
                                      "rn_peer_allreduce_mc_f32)"}.get(collective, "torch.distributed all_reduce, " + collective)))
                              if world > 1 else "none",
                "l2": "per-step working set (%.1f GB of per-ray state on rank 0) is far larger than L2; no flush needed"
                      % (3 * n_rays * M * 4 / 1e9),
            },
            "roofline": {
                "bound": "hbm", "achieved": a_next, "peak": peak, "unit": "GB/s", "frac": a_next / peak,
                "traffic": (traffic or {}).get("bp_kernel_dram_bytes_per_launch"),
                "traffic_source": ((traffic or {}).get("source", "profiles/roofline_traffic.json") +
                                   " (ncu capture of this configuration; not measured by this run)") if traffic else None,
                "kernel": "bp4_kernel<NCH, false>: one NON-FIRST BP sweep over all rays of this rank = one launch per "
                          "ray-length class, timed with CUDA events around the launch set",
                "algorithmic_bytes_per_launch": b_next, "launch_ms": next_ms,
                "launches_timed": int(sweep_ms[:, 1:].size), "peak_source": peak_src,
                "first_sweep": {"kernel": "bp4_first_mapped_kernel<NCH> (plane->voxel mapping + first sweep)" if not args.no_fuse
                                else "bp4_kernel<NCH, true>", "algorithmic_bytes_per_launch": b_first,
                                "launch_ms": first_ms, "achieved": a_first, "frac": a_first / peak},
                "all_sweeps": {"algorithmic_bytes": blend_bytes, "ms": blend_ms, "frac": frac(blend_bytes, blend_ms)},
                "stages": {
                    "frontend": {"algorithmic_bytes": fe_bytes, "ms": fe_ms, "frac": frac(fe_bytes, fe_ms),
                                 "note": "not an HBM-bound stage: every (ray, plane, view) sample is its own 128-byte feature vector, "
                                         "gathered through the L1 (L1 data pipe 65 % busy, issue 58 %; DESIGN.md 4)"},
                    "bp": {"algorithmic_bytes": blend_bytes + grid_bytes, "ms": bp_ms, "frac": frac(blend_bytes + grid_bytes, bp_ms)},
                    "depth": {"algorithmic_bytes": de_bytes, "ms": de_ms, "frac": frac(de_bytes, de_ms)},
                },
                "step_model": {"bytes_per_step_per_gpu": step_bytes, "frac_of_peak": frac(step_bytes, ms_step)},
            },
            "stages_ms": {"frontend": fe_ms, "bp": bp_ms, "depth": de_ms,
                          "bp_first_sweep": first_ms, "bp_next_sweep": next_ms, "exchange_incl_wait_for_slowest_rank": exchange_ms},
            "e2e": e2e, "gpu_launches": int(gpu_launches), "clocks": clocks,
        }
        if cnn is not None:
            line["cnn"] = cnn
        if world == 1 and not args.no_cpu:
            from oracle import oracle as orc
            orc.build()
            orc.set_threads(host_cores())
            sample = cpu_calibrated_sample(cfg, 4.0)
            rays, secs, passes = cpu_timed_passes(sample, 15.0)
            line["cpu_baseline"] = {
                "value": rays / secs, "unit": "rays/s", "cores": orc.num_threads(), "kind": "port",
                "sample": "%d rays of reference image 0 of the same workload (strided pixels), full pipeline "
                          "(front end + %d sweeps + depth), %d passes, %.1f s of CPU work"
                          % (int(sample.ray_idxs.shape[0]), cfg["I"], passes, secs)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"],
                    help="strong: the fixed job sharded by rays over the GPUs (default); weak: fixed work per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the e2e leg (profiling runs only)")
    ap.add_argument("--no-fuse", action="store_true", help="separate plane->voxel mapping kernel instead of the fused first sweep")
    ap.add_argument("--no-balance", action="store_true", help="strong scaling: blocks of equal ray counts instead of equal work")
    ap.add_argument("--collective", default="auto", choices=["auto", "peer", "peer_p2p", "nccl"],
                    help="exchange step of the multi-GPU path (engine.py); auto = peer kernel when the GPUs map each other")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        run_gpu_arm(args, cfg)


if __name__ == "__main__":
    main()
