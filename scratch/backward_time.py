"""Times the training graph from the similarity scores on (raynet_b200.training.forward_backward_pass: plane -> voxel
mapping, I unrolled BP sweeps, depth estimate, loss, and the whole backward pass) on one GPU.  The ray geometry comes
from this library's own tracing kernel.  Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from raynet_b200.engine import RayPotentialEngine
from raynet_b200.synth import SyntheticScene, camera_arrays, get_voxel_grid
from raynet_b200.training import forward_backward_pass

dev = torch.device("cuda", 0)
V, H, W, G, D, M, I = 5, 256, 256, 128, 32, 384, 3
N = int(os.environ.get("RAYS", "20000"))
scene = SyntheticScene(V, H, W, (G, G, G))
eng = RayPotentialEngine(M, D, V, 32, H, W, 11, scene.bbox.ravel(), (G, G, G), max_rays=N)
vgrid = np.ascontiguousarray(get_voxel_grid(scene.bbox.ravel(), np.array([G, G, G], np.int32)).transpose(1, 2, 3, 0))
eng.set_voxel_grid(torch.from_numpy(vgrid).to(dev))
P, P_inv, centre = camera_arrays([scene.get_image(j) for j in scene.view_order(0)])
rng = np.random.default_rng(0)
ids = torch.from_numpy(np.sort(rng.choice(H * W, N, replace=False)).astype(np.int32)).to(dev)
eng.trace_image(ids, torch.from_numpy(P_inv.ravel()).to(dev), torch.from_numpy(centre.ravel()).to(dev))
idx, cnt = eng.voxel_indices(), eng.count[:N].clone()
starts, ends = eng.starts[:N].clone(), eng.ends[:N].clone()
g = torch.Generator(device=dev); g.manual_seed(1)
target = torch.rand((N, M), device=dev, generator=g)
target = target * (torch.arange(M, device=dev)[None, :] < cnt[:, None])
target = target / target.sum(1, keepdim=True).clamp_min(1e-12)
cam = torch.from_numpy(np.tile(np.append(centre.ravel()[:3], 1.0).astype(np.float32), (N, 1))).to(dev)
vg = torch.from_numpy(vgrid).to(dev)
out = {"rays": N, "grid": G, "max_voxels": M, "depth_planes": D, "bp_iterations": I, "mean_voxels_per_ray": float(cnt.float().mean())}
for loss in ("squared_emd", "emd"):
    def step():
        z = torch.randn((N, D), device=dev, generator=g).requires_grad_(True)
        gam = torch.tensor(0.05, device=dev, requires_grad=True)
        L, _ = forward_backward_pass(z, vg, idx, cnt, target, starts, ends, cam, (G, G, G), gamma=gam, bp_iterations=I, loss=loss)
        return L
    for _ in range(3):
        step().backward()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fwd = bwd = 0.0
    reps = 10
    for _ in range(reps):
        e[0].record(); L = step(); e[1].record(); L.backward(); e[2].record(); torch.cuda.synchronize()
        fwd += e[0].elapsed_time(e[1]); bwd += e[1].elapsed_time(e[2])
    out[loss] = {"forward_ms": fwd / reps, "backward_ms": bwd / reps, "rays_per_s_forward_plus_backward": N / ((fwd + bwd) / reps * 1e-3)}
print(json.dumps(out))
