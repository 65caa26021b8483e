"""Multi-GPU check of the ray-sharded forward pass (run under torchrun, one rank per GPU; NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/multi_gpu_check.py

Every rank runs RayNetForwardPass(shard="rays") on the same small scene -- its block of the (image, pixel)
ray enumeration, one all-reduce of the occupancy accumulator per sweep, the depth maps completed on every
rank -- and compares with the same job run alone (shard="none") and, on rank 0, with the CPU oracle.
Also exercised: images in -> SimpleCNN per rank share -> all-gather of the feature maps.
Prints "MULTI_GPU_CHECK ok ..." on rank 0.  tests/test_gpu_multi.py launches it when >= 2 GPUs are visible.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from raynet_b200.common.generation_parameters import GenerationParameters
    from raynet_b200.forward_pass import RayNetForwardPass
    from raynet_b200.models import SimpleCNN
    from raynet_b200.synth import SyntheticScene, random_features

    V, H, W, G, D, M, I = 5, 64, 64, 48, 16, 144, 3
    scene = SyntheticScene(V, H, W, (G, G, G), with_images=True)
    feats = random_features(V, H, W, 32, 11, seed=4) * np.float32(3.0)

    class FeatureModel(object):
        def predict_features(self, scene, views):
            return feats[list(views)]

    gp = GenerationParameters(depth_planes=D, neighbors=V - 1, grid_shape=np.array([G, G, G], np.int32),
                              max_number_of_marched_voxels=M, padding=11, gamma_mrf=0.05)
    out = {}
    for name, model, coll in (("features", FeatureModel(), "auto"), ("features_nccl", FeatureModel(), "nccl"),
                              ("features_p2p", FeatureModel(), "peer_p2p"),
                              ("cnn", SimpleCNN.random_init(channels=3, seed=1), "auto")):
        sharded = RayNetForwardPass(model, gp, "sample_in_bbox", scene.image_shape, H * W, bp_iterations=I, shard="rays",
                                    collective=coll)
        alone = RayNetForwardPass(model, gp, "sample_in_bbox", scene.image_shape, H * W, bp_iterations=I, shard="none")
        a = np.stack(list(sharded.forward_pass(scene, (0, V, 1))))
        b = np.stack(list(alone.forward_pass(scene, (0, V, 1))))
        assert a.shape == b.shape == (V, H, W)
        # float atomics / all-reduce reorder sums: depths agree except where the arg-max is a near tie
        same = np.abs(a - b) < 1e-6
        occ_a = sharded.engine.occupancy().cpu().numpy()
        occ_b = alone.engine.occupancy().cpu().numpy()
        out[name] = (float(same.mean()), float(np.abs(occ_a - occ_b).max()), sharded.engine.n_rays, alone.engine.n_rays,
                     sharded.engine.collective)
        # twice more on the same engine: the flags / epochs of the exchange kernel carry over between calls
        for _ in range(2):
            a2 = np.stack(list(sharded.forward_pass(scene, (0, V, 1))))
            assert (np.abs(a2 - b) < 1e-6).mean() > 0.999
        assert same.mean() > 0.999, (name, same.mean())
        assert np.abs(occ_a - occ_b).max() <= 1e-5, (name, np.abs(occ_a - occ_b).max())
        assert sharded.engine.n_rays < alone.engine.n_rays or world == 1
        # every rank holds the same complete maps
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi)
    if rank == 0:
        from oracle import oracle as orc
        from raynet_b200.synth import camera_arrays, get_voxel_grid
        orc.build()
        bbox = scene.bbox.ravel()
        grid = np.array([G, G, G], np.int32)
        vgrid = np.ascontiguousarray(get_voxel_grid(bbox, grid).transpose(1, 2, 3, 0))
        ids = np.arange(H * W, dtype=np.int32)
        fronts = []
        for i in range(V):
            order = scene.view_order(i)
            P, P_inv, centre = camera_arrays([scene.get_image(j) for j in order])
            fronts.append(orc.frontend(ids, np.ascontiguousarray(feats[order]), P, P_inv, centre, vgrid, bbox, grid, M, D, V,
                                       32, H, W, 11))
        idx = np.concatenate([f["idx"] for f in fronts])
        cnt = np.concatenate([f["cnt"] for f in fronts])
        S_vox = np.concatenate([f["S_vox"] for f in fronts])
        acc, _ = orc.belief_propagation(S_vox, idx, cnt, grid, gamma=0.05, bp_iterations=I, acc_f64=True)
        sharded = RayNetForwardPass(FeatureModel(), gp, "sample_in_bbox", scene.image_shape, H * W, bp_iterations=I,
                                    shard="none")
        list(sharded.forward_pass(scene, (0, V, 1)))
        err = float(np.abs(sharded.engine.occupancy().cpu().numpy() - orc.occupancy(acc)).max())
        assert err <= 1e-5, err
        print("MULTI_GPU_CHECK ok world=%d %r oracle_err=%.2e" % (world, out, err), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        print("MULTI_GPU_CHECK failed on rank %s\n%s" % (os.environ.get("RANK"), traceback.format_exc()), flush=True)
        raise
