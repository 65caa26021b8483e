"""Launch the reference's own CUDA kernels (cubins built by oracle/build_ref_cuda.py) on torch tensors.

TEST INFRASTRUCTURE ONLY -- a GPU-side oracle and speed bar.  The product path
never imports this module.  The launch closures mirror raynet_fp / raynet_de of the reference
(cuda_implementations/raynet_fp.py:275-376): same argument order, caller-allocated outputs, one
thread per ray; `threads` defaults to 256 (the reference's default of 2048 is not a legal block size).
"""
import ctypes
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
CUBIN_DIR = os.path.join(HERE, "_ref", "cuda")


def available(name):
    return os.path.exists(os.path.join(CUBIN_DIR, "raynet_fp_%s.cubin" % name))


def _check(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError("CUDA driver error %s" % (err,))
    return res[1] if len(res) == 2 else res[1:]


class RefCuda(object):
    """The module `perform_raynet_fp(M, D, N, F, H, W, padding, bbox, grid_shape, "sample_in_bbox")` would JIT."""

    def __init__(self, name):
        import torch
        from cuda.bindings import driver
        self.driver = driver
        torch.cuda.init()
        torch.zeros(1, device="cuda")      # makes the primary context current
        self.params = json.load(open(os.path.join(CUBIN_DIR, "raynet_fp_%s.json" % name)))
        with open(os.path.join(CUBIN_DIR, "raynet_fp_%s.cubin" % name), "rb") as f:
            image = f.read()
        self.module = _check(driver.cuModuleLoadData(image))
        self.fp = _check(driver.cuModuleGetFunction(self.module, b"batch_raynet_fp"))
        self.de = _check(driver.cuModuleGetFunction(self.module, b"batch_complete_depth_estimation"))

    def _launch(self, fn, n_rays, tensors, threads):
        import torch
        for t in tensors:
            assert t.is_cuda and t.is_contiguous()
        blocks = (n_rays + threads - 1) // threads
        values = (int(n_rays),) + tuple(int(t.data_ptr()) for t in tensors)
        types = (ctypes.c_int,) + (ctypes.c_void_p,) * len(tensors)
        stream = torch.cuda.current_stream().cuda_stream
        _check(self.driver.cuLaunchKernel(fn, blocks, 1, 1, threads, 1, 1, 0, stream, (values, types), 0))

    def raynet_fp(self, ray_idxs, features, P, P_inv, camera_center, voxel_grid, ray_voxel_indices, ray_voxel_count,
                  S_voxel_space, acc_in, msgs, acc_out, threads=256):
        """raynet_fp.py:275-326: front end + one BP sweep; msgs updated in place (passed as input and output)."""
        n = int(S_voxel_space.shape[0])
        self._launch(self.fp, n, [ray_idxs, features, P, P_inv, camera_center, voxel_grid, ray_voxel_indices,
                                  ray_voxel_count, S_voxel_space, acc_in, msgs, acc_out, msgs], threads)
        return msgs

    def raynet_de(self, ray_idxs, features, P, P_inv, camera_center, voxel_grid, ray_voxel_indices, ray_voxel_count,
                  S_voxel_space, acc, msgs, depth_map, threads=256):
        """raynet_fp.py:329-376: front end + depth re-estimation + arg-max -> depth."""
        n = int(S_voxel_space.shape[0])
        self._launch(self.de, n, [ray_idxs, features, P, P_inv, camera_center, voxel_grid, ray_voxel_indices,
                                  ray_voxel_count, S_voxel_space, acc, msgs, depth_map], threads)
        return depth_map
