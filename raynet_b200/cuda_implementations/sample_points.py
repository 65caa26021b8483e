"""Drop-ins for raynet/cuda_implementations/sample_points.py."""
import numpy as np

from .. import _lib
from .raynet_fp import _check_scheme
from .utils import all_arrays_to_gpu, current_stream_ptr, ptr, to_gpu


def batch_sample_points(D, H, W, bbox, sampling_scheme):
    """sample_points.py:12-54: sp(ray_idxs, P_inv, camera_center, points f32[B, D, 4])."""
    _check_scheme(sampling_scheme)
    params = _lib.make_params(0, D, 0, 0, H, W, 0, bbox, None)

    @all_arrays_to_gpu
    def sp(ray_idxs, P_inv, camera_center, points, threads=2048):
        n_rays = len(ray_idxs)
        assert points.shape[0] >= n_rays and points.shape[1:] == (D, 4)
        _lib.call("rn_sample_points", params, ptr(ray_idxs), ptr(P_inv), ptr(camera_center), ptr(points),
                  n_rays, current_stream_ptr())

    return sp


def sample_points(ray_idxs, P_inv, camera_center, points, H, W, bbox, sampling_scheme="sample_in_bbox",
                  batch_size=80000):
    """sample_points.py:57-91."""
    _, D, _ = points.shape
    assert points.shape == (H * W, D, 4)
    ray_idxs = to_gpu(ray_idxs.astype(np.int32))
    P_inv_gpu = to_gpu(np.asarray(P_inv, dtype=np.float32).ravel())
    camera_center_gpu = to_gpu(np.asarray(camera_center, dtype=np.float32))
    points_gpu = to_gpu(np.zeros((batch_size, D, 4), dtype=np.float32))
    sp = batch_sample_points(D, H, W, bbox, sampling_scheme)
    for i in range(0, len(ray_idxs), batch_size):
        points_gpu.fill(0)
        sp(ray_idxs[i:i + batch_size], P_inv_gpu, camera_center_gpu, points_gpu)
        n = min(batch_size, len(ray_idxs) - i)
        points[i:i + n, :, :] = points_gpu.get()[:n]
    return points


def compute_depth_from_distribution(ray_idxs, P_inv, camera_center, H, W, bbox, S, depth_map,
                                    sampling_scheme="sample_in_bbox", batch_size=80000):
    """sample_points.py:94-132: depth of the arg-max plane of S for every ray."""
    _, D = S.shape
    ray_idxs_gpu = to_gpu(ray_idxs.astype(np.int32))
    P_inv_gpu = to_gpu(np.asarray(P_inv, dtype=np.float32).ravel())
    camera_center = np.asarray(camera_center, dtype=np.float32)
    camera_center_gpu = to_gpu(camera_center)
    points_gpu = to_gpu(np.zeros((batch_size, D, 4), dtype=np.float32))
    sp = batch_sample_points(D, H, W, bbox, sampling_scheme)
    for i in range(0, len(ray_idxs), batch_size):
        points_gpu.fill(0)
        sp(ray_idxs_gpu[i:i + batch_size], P_inv_gpu, camera_center_gpu, points_gpu)
        idxs = ray_idxs[i:i + batch_size]
        pts = points_gpu.get()[:len(idxs)].transpose(2, 0, 1)
        pts = pts[:-1, np.arange(len(idxs)), S[idxs].argmax(axis=-1)]
        depth_map[idxs] = np.sqrt(np.sum((camera_center.reshape(4, 1)[:-1] - pts) ** 2, axis=0))
    return depth_map
