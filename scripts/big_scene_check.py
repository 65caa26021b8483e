#!/usr/bin/env python
"""Bounded-memory run of the forward-pass surface on a scene that does NOT fit resident:
24 reference images of 1280 x 720 pixels, 256 x 256 x 128 grid, M = 650 (the reference's defaults,
scripts/arguments.py), 4 neighbours, 32 planes, 3 sweeps -- 22.1 M rays whose resident state would be 204 GB.
RayNetForwardPass keeps as many images resident as the budget allows and streams the rest (their voxel-space
rows are recomputed on every sweep, engine.py).  Prints one JSON line; run under gpurun, one GPU.

    python scripts/big_scene_check.py [n_images] [budget_GB]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from raynet_b200.common.generation_parameters import GenerationParameters
    from raynet_b200.forward_pass import RayNetForwardPass
    from raynet_b200.synth import SyntheticScene
    n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    budget = float(sys.argv[2]) * 1e9 if len(sys.argv) > 2 else None
    H, W, G, D, M, V, I = 720, 1280, (256, 256, 128), 32, 650, 5, 3
    scene = SyntheticScene(n_images, H, W, G, neighbors=V - 1)
    F, pad = 32, 11
    g = torch.Generator(device="cpu")
    g.manual_seed(0)
    feats = torch.randn((n_images, H + pad + 1, W + pad + 1, F), generator=g, dtype=torch.float32) / float(np.sqrt(F))
    feats[:, 0] = 0
    feats[:, :, 0] = 0
    feats = feats.cuda()

    class Model(object):
        def predict_features(self, scene, views):
            return feats[list(views)] if list(views) != list(range(n_images)) else feats

    gp = GenerationParameters(depth_planes=D, neighbors=V - 1, grid_shape=np.array(G, np.int32),
                              max_number_of_marched_voxels=M, padding=pad, gamma_mrf=0.05)
    fp = RayNetForwardPass(Model(), gp, "sample_in_bbox", scene.image_shape, rays_batch=50000, bp_iterations=I,
                           memory_budget=budget)
    free0, total = torch.cuda.mem_get_info()
    t0 = time.perf_counter()
    maps = list(fp.forward_pass(scene, (0, n_images, 1)))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    maps2 = list(fp.forward_pass(scene, (0, n_images, 1)))
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    eng = fp.engine
    resident = [eng.is_resident(k) for k in range(len(eng.segments))]
    out = {
        "scene": "%d images %dx%d, grid %s, M=%d, %d views, %d planes, %d sweeps" % (n_images, W, H, "x".join(map(str, G)), M, V, D, I),
        "rays": eng.n_rays, "state_if_resident_GB": eng.n_rays * eng.bytes_per_ray(True) / 1e9,
        "memory_budget_GB": eng.memory_budget / 1e9, "hbm_total_GB": total / 1e9,
        "segments_resident": int(sum(resident)), "segments_streamed": int(len(resident) - sum(resident)),
        "first_call_s": t1 - t0, "second_call_s": t2 - t1, "rays_per_s": eng.n_rays / (t2 - t1),
        "peak_allocated_GB": torch.cuda.max_memory_allocated() / 1e9,
        "maps": len(maps), "finite": bool(all(np.isfinite(m).all() for m in maps)),
        "repeatable": float(np.mean([np.mean(np.abs(a - b) < 1e-6) for a, b in zip(maps, maps2)])),
    }
    # a job that cannot fit even streamed must fail with a clear error before anything is launched
    try:
        RayNetForwardPass(Model(), gp, "sample_in_bbox", scene.image_shape, rays_batch=50000, bp_iterations=I,
                          memory_budget=8e9).forward_pass(scene, (0, n_images, 1)).__next__()
        out["tiny_budget"] = "ran (unexpected)"
    except MemoryError as e:
        out["tiny_budget"] = "MemoryError: " + str(e)[:160]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
