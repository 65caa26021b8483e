"""Drop-ins for raynet/cuda_implementations/mvcnn_with_ray_marching_and_voxels_mapping.py:
similarity + DDA + plane->voxel mapping without the MRF."""
import numpy as np

from .. import _lib
from .raynet_fp import _check_scheme
from .utils import all_arrays_to_gpu, current_stream_ptr, ptr


def batch_mvcnn_voxel_traversal_with_ray_marching(M, D, N, F, H, W, padding, bbox, grid_shape,
                                                  sampling_scheme):
    """:11-174.  Returns mvcnn_voxels(ray_idxs, features, P, P_inv, camera_center, voxel_grid,
    ray_voxel_indices, ray_voxel_count, S_new)."""
    _check_scheme(sampling_scheme)
    grid_shape = tuple(int(g) for g in np.asarray(grid_shape).ravel())
    params = _lib.make_params(M, D, N, F, H, W, padding, bbox, grid_shape)

    @all_arrays_to_gpu
    def mvcnn_voxels(ray_idxs, features, P, P_inv, camera_center, voxel_grid, ray_voxel_indices,
                     ray_voxel_count, S_new, threads=2048):
        assert S_new.shape[1] == M
        assert ray_voxel_indices.shape[1:] == (M, 3)
        assert len(ray_voxel_count.shape) == 1
        assert np.float32 == S_new.dtype
        assert np.int32 == ray_voxel_indices.dtype
        assert np.int32 == ray_voxel_count.dtype
        n_rays = min(len(S_new), len(ray_idxs))
        _lib.call("rn_mvcnn_voxel", params, ptr(ray_idxs), ptr(features), ptr(P), ptr(P_inv),
                  ptr(camera_center), ptr(voxel_grid), ptr(ray_voxel_indices), ptr(ray_voxel_count),
                  ptr(S_new), n_rays, current_stream_ptr())
        return S_new

    return mvcnn_voxels


def batch_mvcnn_voxel_traversal_with_ray_marching_with_depth_estimation(M, D, N, F, H, W, padding, bbox,
                                                                        grid_shape, sampling_scheme):
    """:177-377.  Additionally fills depth_map with the distance of the arg-max voxel centre."""
    _check_scheme(sampling_scheme)
    grid_shape = tuple(int(g) for g in np.asarray(grid_shape).ravel())
    params = _lib.make_params(M, D, N, F, H, W, padding, bbox, grid_shape)

    @all_arrays_to_gpu
    def mvcnn_voxels_depth(ray_idxs, features, P, P_inv, camera_center, voxel_grid, ray_voxel_indices,
                           ray_voxel_count, S_new, depth_map, threads=2048):
        assert S_new.shape[1] == M
        assert ray_voxel_indices.shape[1:] == (M, 3)
        assert len(ray_voxel_count.shape) == 1
        assert np.float32 == S_new.dtype
        assert np.int32 == ray_voxel_indices.dtype
        assert np.int32 == ray_voxel_count.dtype
        n_rays = min(len(S_new), len(ray_idxs), len(depth_map))
        _lib.call("rn_mvcnn_voxel_depth", params, ptr(ray_idxs), ptr(features), ptr(P), ptr(P_inv),
                  ptr(camera_center), ptr(voxel_grid), ptr(ray_voxel_indices), ptr(ray_voxel_count),
                  ptr(S_new), ptr(depth_map), n_rays, current_stream_ptr())
        return depth_map

    return mvcnn_voxels_depth
