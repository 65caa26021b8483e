"""Drop-ins for raynet/ray_marching/ray_tracing_cuda.py."""
import numpy as np

from .. import _lib
from ..cuda_implementations.utils import all_arrays_to_gpu, current_stream_ptr, ptr, to_gpu


def batch_voxel_traversal(M, bbox, grid_shape):
    """ray_tracing_cuda.py:12-63: vtr(points_start, points_end, ray_voxel_indices, ray_voxel_count)."""
    grid_shape = tuple(int(g) for g in np.asarray(grid_shape).ravel())
    params = _lib.make_params(M=M, bbox=bbox, grid_shape=grid_shape)

    @all_arrays_to_gpu
    def vtr(points_start, points_end, ray_voxel_indices, ray_voxel_count, threads=1024):
        assert ray_voxel_indices.shape[1] == M
        assert np.float32 == points_start.dtype
        assert np.float32 == points_end.dtype
        assert np.int32 == ray_voxel_indices.dtype
        assert np.int32 == ray_voxel_count.dtype
        n_rays = len(ray_voxel_count)
        _lib.call("rn_voxel_traversal", params, ptr(points_start), ptr(points_end), ptr(ray_voxel_indices),
                  ptr(ray_voxel_count), n_rays, current_stream_ptr())

    return vtr


def voxel_traversal(bbox, grid_shape, ray_voxel_indices, ray_start, ray_end):
    """ray_tracing_cuda.py:66-89: single ray, same signature as the Cython voxel_traversal
    (ray_tracing.pyx:64); fills ray_voxel_indices (N, 3) in place and returns the count."""
    N, _ = ray_voxel_indices.shape
    ray_voxel_count = to_gpu(np.zeros((1,), dtype=np.int32))
    ray_voxel_indices_out = to_gpu(np.ascontiguousarray(ray_voxel_indices).reshape(1, N, 3))
    vtr = batch_voxel_traversal(N, bbox, grid_shape)
    vtr(to_gpu(np.asarray(ray_start, dtype=np.float32).reshape(1, 3)),
        to_gpu(np.asarray(ray_end, dtype=np.float32).reshape(1, 3)),
        ray_voxel_indices_out, ray_voxel_count, threads=1)
    ray_voxel_indices[:, :] = ray_voxel_indices_out.get()[0]
    return int(ray_voxel_count.get()[0])


def perform_ray_marching(scene, img_idx, M, rays_idxs, grid_shape, batch_size=40000):
    """ray_tracing_cuda.py:92-143.  The reference gets the ray/bbox intersections from a TF
    graph (tf_implementations/sampling_schemes.py:151-172); here they come from the same
    sample_in_bbox the fused kernels use (rn_sample_in_bbox)."""
    H, W = scene.image_shape
    ref_camera = scene.get_image(img_idx).camera
    N = rays_idxs.shape[0]
    params = _lib.make_params(H=H, W=W, bbox=np.asarray(scene.bbox).ravel())
    ray_idxs_gpu = to_gpu(np.asarray(rays_idxs, dtype=np.int32))
    starts = to_gpu(np.zeros((N, 3), dtype=np.float32))
    ends = to_gpu(np.zeros((N, 3), dtype=np.float32))
    P_inv = to_gpu(np.asarray(ref_camera.P_pinv, dtype=np.float32).ravel())
    centre = to_gpu(np.asarray(ref_camera.center, dtype=np.float32).ravel())
    _lib.call("rn_sample_in_bbox", params, ptr(ray_idxs_gpu), ptr(P_inv), ptr(centre), ptr(starts), ptr(ends),
              N, current_stream_ptr())
    ray_voxel_indices = np.zeros((N, M, 3), dtype=np.int32)
    ray_voxel_count = to_gpu(np.zeros((N,), dtype=np.int32))
    ray_voxel_indices_gpu = to_gpu(np.zeros((batch_size, M, 3), dtype=np.int32))
    vtr = batch_voxel_traversal(M, np.asarray(scene.bbox).ravel(), grid_shape)
    for r in range(0, N, batch_size):
        ray_voxel_indices_gpu.fill(0)
        n = min(batch_size, N - r)
        vtr(starts[r:r + n], ends[r:r + n], ray_voxel_indices_gpu, ray_voxel_count[r:r + n])
        ray_voxel_indices[r:r + n, :, :] = ray_voxel_indices_gpu.get()[:n]
    return ray_voxel_indices, ray_voxel_count.get()
