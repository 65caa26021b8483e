"""Drop-ins for raynet/planes_voxels_mapping/planes_voxels_mapping_cuda.py."""
import numpy as np

from .. import _lib
from ..cuda_implementations.utils import all_arrays_to_gpu, current_stream_ptr, ptr, to_gpu


def batch_depth_to_voxels_mapping(M, D, grid_shape):
    """planes_voxels_mapping_cuda.py:11-67: pvm(voxel_grid, ray_voxel_indices, ray_voxel_count,
    ray_start, ray_end, S, S_new) -> S_new."""
    grid_shape = tuple(int(g) for g in np.asarray(grid_shape).ravel())
    params = _lib.make_params(M=M, D=D, grid_shape=grid_shape)

    @all_arrays_to_gpu
    def pvm(voxel_grid, ray_voxel_indices, ray_voxel_count, ray_start, ray_end, S, S_new, threads=2048):
        assert S.shape[1] == D
        assert S_new.shape[1] == M
        assert len(ray_voxel_count.shape) == 1
        assert np.float32 == S.dtype
        assert np.float32 == S_new.dtype
        assert np.int32 == ray_voxel_count.dtype
        assert np.float32 == ray_start.dtype
        assert np.float32 == ray_end.dtype
        _lib.call("rn_planes_to_voxels", params, ptr(voxel_grid), ptr(ray_voxel_indices), ptr(ray_voxel_count),
                  ptr(ray_start), ptr(ray_end), ptr(S), ptr(S_new), len(S), current_stream_ptr())
        return S_new

    return pvm


def depth_to_voxels(ray_voxel_count, ray_voxel_indices, rays_idxs, voxel_grid, points, S, S_new,
                    batch_size=20000):
    """planes_voxels_mapping_cuda.py:70-124.  voxel_grid (3, Gx, Gy, Gz), points (4, N, D)."""
    N, M, _ = ray_voxel_indices.shape
    _, _, D = points.shape
    S_new.fill(0)
    points_start_gpu = to_gpu(np.ascontiguousarray(points[:-1, rays_idxs, 0].T, dtype=np.float32))
    points_end_gpu = to_gpu(np.ascontiguousarray(points[:-1, rays_idxs, -1].T, dtype=np.float32))
    ray_voxel_count_gpu = to_gpu(np.ascontiguousarray(ray_voxel_count[rays_idxs], dtype=np.int32))
    pvm = batch_depth_to_voxels_mapping(M, D, np.array(voxel_grid.shape[1:]))
    voxel_grid_gpu = to_gpu(np.ascontiguousarray(voxel_grid.transpose(1, 2, 3, 0), dtype=np.float32).ravel())
    for i in range(0, len(rays_idxs), batch_size):
        sel = rays_idxs[i:i + batch_size]
        s = pvm(voxel_grid_gpu, np.ascontiguousarray(ray_voxel_indices[sel]),
                ray_voxel_count_gpu[i:i + batch_size], points_start_gpu[i:i + batch_size],
                points_end_gpu[i:i + batch_size], np.ascontiguousarray(S[sel], dtype=np.float32),
                np.ascontiguousarray(S_new[sel]))
        S_new[sel] = s.get()
    return S_new
