import sys, time, os, numpy as np, torch
sys.path.insert(0, '.')
import bench
from raynet_b200.common.generation_parameters import GenerationParameters
from raynet_b200.forward_pass import RayNetForwardPass
cfg = bench.CONFIGS['c3']
H, W, G, V, D, M, I = (cfg[k] for k in ("H", "W", "G", "V", "D", "M", "I"))
scene = bench.make_scene(cfg, 1)
views = sorted(set(v for i in range(scene.n_images) for v in scene.view_order(i)))
model = bench.FeatureModel(views, H, W)
gp = GenerationParameters(depth_planes=D, neighbors=V - 1, grid_shape=np.array([G, G, G], np.int32),
                          max_number_of_marched_voxels=M, padding=bench.PADDING, gamma_mrf=bench.GAMMA)
fp = RayNetForwardPass(model, gp, "sample_in_bbox", scene.image_shape, rays_batch=H * W, bp_iterations=I)
rng = (0, scene.n_images, 1)
for _ in range(3):
    list(fp.forward_pass(scene, rng))
torch.cuda.synchronize()
# host-side profile of one call
import cProfile, pstats
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
maps = list(fp.forward_pass(scene, rng))
pr.disable()
torch.cuda.synchronize()
print('e2e step %.2f ms' % ((time.perf_counter() - t0) * 1e3))
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
# pure H2D timing
f = model.host
dev = torch.empty(f.shape, dtype=torch.float32, device='cuda')
torch.cuda.synchronize(); t0 = time.perf_counter(); dev.copy_(f, non_blocking=True); torch.cuda.synchronize()
dt = time.perf_counter() - t0
print('H2D %.1f MB in %.2f ms = %.1f GB/s, pinned=%s' % (f.numel() * 4 / 1e6, dt * 1e3, f.numel() * 4 / dt / 1e9, f.is_pinned()))
