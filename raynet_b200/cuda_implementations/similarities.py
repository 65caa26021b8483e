"""Drop-ins for raynet/cuda_implementations/similarities.py (plane-sweep similarity only)."""
import numpy as np

from .. import _lib
from .raynet_fp import _check_scheme
from .utils import all_arrays_to_gpu, current_stream_ptr, ptr, to_gpu


def perform_multi_view_cnn_forward_pass(D, N, F, H, W, padding, bbox, sampling_scheme):
    """similarities.py:11-130: returns mvcnnfp(ray_idxs, features, P, P_inv, camera_center, S)."""
    _check_scheme(sampling_scheme)
    params = _lib.make_params(0, D, N, F, H, W, padding, bbox, None)

    @all_arrays_to_gpu
    def mvcnnfp(ray_idxs, features, P, P_inv, camera_center, S, threads=2048):
        assert S.shape[1] == D
        assert np.float32 == S.dtype
        n_rays = min(len(S), len(ray_idxs))
        _lib.call("rn_mvcnn_forward", params, ptr(ray_idxs), ptr(features), ptr(P), ptr(P_inv),
                  ptr(camera_center), ptr(S), n_rays, current_stream_ptr())
        return S

    return mvcnnfp


def perform_multi_view_cnn_forward_pass_with_depth_estimation(D, N, F, H, W, padding, bbox, sampling_scheme):
    """similarities.py:133-285: returns mvcnnfp(ray_idxs, features, P, P_inv, camera_center, S, points,
    depth_map)."""
    _check_scheme(sampling_scheme)
    params = _lib.make_params(0, D, N, F, H, W, padding, bbox, None)

    @all_arrays_to_gpu
    def mvcnnfp(ray_idxs, features, P, P_inv, camera_center, S, points, depth_map, threads=2048):
        assert S.shape[1] == D
        assert np.float32 == S.dtype
        n_rays = min(len(S), len(ray_idxs), len(depth_map))
        _lib.call("rn_mvcnn_forward_depth", params, ptr(ray_idxs), ptr(features), ptr(P), ptr(P_inv),
                  ptr(camera_center), ptr(S), ptr(points), ptr(depth_map), n_rays, current_stream_ptr())
        return depth_map

    return mvcnnfp


def multi_view_cnn_fp(ray_idxs, features, P, P_inv, camera_center, bbox, S, padding, sampling_scheme,
                      batch_size=80000):
    """similarities.py:288-341: batched host driver; S (N, D) is filled in place (numpy) and
    returned."""
    N, D = S.shape
    n_views, Hp, Wp, F = features.shape
    H, W = Hp - padding - 1, Wp - padding - 1
    fp = perform_multi_view_cnn_forward_pass(D, n_views, F, H, W, padding, np.asarray(bbox).ravel(),
                                             sampling_scheme)
    features_gpu = to_gpu(np.ascontiguousarray(features, dtype=np.float32).ravel())
    P_gpu = to_gpu(np.asarray(P, dtype=np.float32).ravel())
    P_inv_gpu = to_gpu(np.asarray(P_inv, dtype=np.float32).ravel())
    c_gpu = to_gpu(np.asarray(camera_center, dtype=np.float32).ravel())
    ray_idxs_gpu = to_gpu(np.asarray(ray_idxs, dtype=np.int32))
    for i in range(0, N, batch_size):
        s_gpu = to_gpu(np.zeros((min(batch_size, N - i), D), dtype=np.float32))
        fp(ray_idxs_gpu[i:i + batch_size], features_gpu, P_gpu, P_inv_gpu, c_gpu, s_gpu)
        S[i:i + batch_size] = s_gpu.get()
    return S
