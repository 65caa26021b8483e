"""Per-stage host wall clock of RayNetForwardPass.forward_pass on C3 (profile=True: the device is drained at every
stamp, so the stages do not overlap -- the sum is an upper bound of the e2e time).  Run alone or under torchrun."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from raynet_b200.common.generation_parameters import GenerationParameters
from raynet_b200.forward_pass import RayNetForwardPass
from raynet_b200.models import SimpleCNN
cfg = bench.CONFIGS["c3"]
H, W, G, V, D, M, I = (cfg[k] for k in ("H", "W", "G", "V", "D", "M", "I"))
scene = bench.make_scene(cfg, world, "strong", with_images=True)
gp = GenerationParameters(depth_planes=D, neighbors=V - 1, grid_shape=np.array([G, G, G], np.int32),
                          max_number_of_marched_voxels=M, padding=11, gamma_mrf=0.05)
fp = RayNetForwardPass(SimpleCNN.random_init(channels=3, seed=0), gp, "sample_in_bbox", scene.image_shape, H * W, bp_iterations=I)
for _ in range(3):
    list(fp.forward_pass(scene, (0, scene.n_images, 1)))
fp.profile = True
acc = {}
for _ in range(5):
    list(fp.forward_pass(scene, (0, scene.n_images, 1)))
    for k, v in fp.timings.items():
        acc[k] = acc.get(k, 0.0) + v / 5
if rank == 0:
    print(json.dumps({"world": world, "ms": {k: round(v, 3) for k, v in acc.items()}, "sum": round(sum(acc.values()), 3)}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
