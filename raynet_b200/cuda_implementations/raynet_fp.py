"""perform_raynet_fp -- drop-in for raynet/cuda_implementations/raynet_fp.py:10-378.

Same factory arguments, same closure signatures, same in-place / return behaviour; the
work is done by rn_raynet_fp / rn_raynet_de of the C-ABI library (no JIT: the sizes the
reference substitutes into its CUDA source travel in an RnParams struct).
"""
import numpy as np

from .. import _lib
from .utils import all_arrays_to_gpu, current_stream_ptr, ptr

SUPPORTED_SAMPLING_SCHEMES = ("sample_in_bbox",)


def _check_scheme(sampling_scheme):
    if sampling_scheme not in SUPPORTED_SAMPLING_SCHEMES:
        # sampling_schemes.cu defines only sample_in_bbox; anything else fails to compile
        # in the reference (nvcc error from SourceModule)
        raise NotImplementedError("sampling scheme %r is not implemented" % (sampling_scheme,))


def perform_raynet_fp(M, D, N, F, H, W, padding, bbox, grid_shape, sampling_scheme):
    """Arguments as raynet_fp.py:22-41.  Returns (raynet_fp, raynet_de)."""
    _check_scheme(sampling_scheme)
    grid_shape = tuple(int(g) for g in np.asarray(grid_shape).ravel())
    params = _lib.make_params(M, D, N, F, H, W, padding, bbox, grid_shape)

    def _common_asserts(S_voxel_space, ray_voxel_indices, ray_voxel_count, msgs, acc):
        # raynet_fp.py:290-301 / :344-354
        assert S_voxel_space.shape[1] == M
        assert ray_voxel_indices.shape[1:] == (M, 3)
        assert len(ray_voxel_count.shape) == 1
        assert len(ray_voxel_count) == len(S_voxel_space) == len(ray_voxel_indices)
        assert S_voxel_space.shape[1] == msgs.shape[1]
        assert acc.shape == tuple(grid_shape)
        assert np.float32 == S_voxel_space.dtype
        assert np.float32 == msgs.dtype
        assert np.int32 == ray_voxel_indices.dtype
        assert np.int32 == ray_voxel_count.dtype

    @all_arrays_to_gpu
    def raynet_fp(ray_idxs, features, P, P_inv, camera_center, voxel_grid, ray_voxel_indices,
                  ray_voxel_count, S_voxel_space, ray_to_occupancy_accumulated_pon,
                  ray_to_occupancy_messages_pon, ray_to_occupancy_accumulated_out_pon, threads=2048):
        _common_asserts(S_voxel_space, ray_voxel_indices, ray_voxel_count,
                        ray_to_occupancy_messages_pon, ray_to_occupancy_accumulated_pon)
        assert ray_to_occupancy_accumulated_out_pon.shape == tuple(grid_shape)
        # the reference launches one thread per row of S_voxel_space; rows beyond the ray
        # list would read past ray_idxs, so only min(len) rays are meaningful
        n_rays = min(len(S_voxel_space), len(ray_idxs), len(ray_to_occupancy_messages_pon))
        _lib.call("rn_raynet_fp", params, ptr(ray_idxs), ptr(features), ptr(P), ptr(P_inv), ptr(camera_center),
                  ptr(voxel_grid), ptr(ray_voxel_indices), ptr(ray_voxel_count), ptr(S_voxel_space),
                  ptr(ray_to_occupancy_accumulated_pon), ptr(ray_to_occupancy_messages_pon),
                  ptr(ray_to_occupancy_accumulated_out_pon), n_rays, current_stream_ptr())
        return ray_to_occupancy_messages_pon

    @all_arrays_to_gpu
    def raynet_de(ray_idxs, features, P, P_inv, camera_center, voxel_grid, ray_voxel_indices,
                  ray_voxel_count, S_voxel_space, ray_to_occupancy_accumulated_pon,
                  ray_to_occupancy_messages_pon, depth_map, threads=2048):
        _common_asserts(S_voxel_space, ray_voxel_indices, ray_voxel_count,
                        ray_to_occupancy_messages_pon, ray_to_occupancy_accumulated_pon)
        n_rays = min(len(S_voxel_space), len(ray_idxs), len(ray_to_occupancy_messages_pon), len(depth_map))
        _lib.call("rn_raynet_de", params, ptr(ray_idxs), ptr(features), ptr(P), ptr(P_inv), ptr(camera_center),
                  ptr(voxel_grid), ptr(ray_voxel_indices), ptr(ray_voxel_count), ptr(S_voxel_space),
                  ptr(ray_to_occupancy_accumulated_pon), ptr(ray_to_occupancy_messages_pon), ptr(depth_map),
                  n_rays, current_stream_ptr())
        return depth_map

    return raynet_fp, raynet_de
