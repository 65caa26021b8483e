"""Drop-ins for raynet/mrf/mrf_cuda.py: ray-potential BP on precomputed voxel lists."""
import numpy as np

from .. import _lib
from ..cuda_implementations.utils import all_arrays_to_gpu, current_stream_ptr, ptr, to_gpu


def batch_ray_belief_propagation(M, grid_shape, parity=False):
    """mrf_cuda.py:12-124.  Returns (bp, de).  parity=True: the accumulators are float64 device arrays
    and the kernels follow mrf_np.py as it executes under NumPy >= 2 (csrc/rn_parity.cuh)."""
    grid_shape = tuple(int(g) for g in np.asarray(grid_shape).ravel())
    params = _lib.make_params(M=M, grid_shape=grid_shape)
    acc_dtype = np.float64 if parity else np.float32
    suffix = "_f64" if parity else ""

    @all_arrays_to_gpu
    def bp(S, ray_voxel_indices, ray_voxel_count, ray_to_occupancy_accumulated_pon,
           ray_to_occupancy_messages_pon, ray_to_occupancy_accumulated_out_pon, threads=1024):
        # mrf_cuda.py:47-59
        assert S.shape[1] == M
        assert ray_voxel_indices.shape[1:] == (M, 3)
        assert len(ray_voxel_count.shape) == 1
        assert len(ray_voxel_count) == len(S) == len(ray_voxel_indices)
        assert len(ray_voxel_count) == len(ray_to_occupancy_messages_pon)
        assert S.shape[1] == ray_to_occupancy_messages_pon.shape[1]
        assert ray_to_occupancy_accumulated_pon.shape == tuple(grid_shape)
        assert ray_to_occupancy_accumulated_out_pon.shape == tuple(grid_shape)
        assert np.float32 == S.dtype
        assert np.float32 == ray_to_occupancy_messages_pon.dtype
        assert np.int32 == ray_voxel_indices.dtype
        assert np.int32 == ray_voxel_count.dtype
        assert acc_dtype == ray_to_occupancy_accumulated_pon.dtype == ray_to_occupancy_accumulated_out_pon.dtype
        _lib.call("rn_bp_iteration" + suffix, params, ptr(S), ptr(ray_voxel_indices), ptr(ray_voxel_count),
                  ptr(ray_to_occupancy_accumulated_pon), ptr(ray_to_occupancy_messages_pon),
                  ptr(ray_to_occupancy_accumulated_out_pon), len(S), current_stream_ptr())
        return ray_to_occupancy_accumulated_out_pon, ray_to_occupancy_messages_pon

    @all_arrays_to_gpu
    def de(S, ray_voxel_indices, ray_voxel_count, ray_to_occupancy_accumulated_pon,
           ray_to_occupancy_messages_pon, S_new, threads=1024):
        # mrf_cuda.py:91-104
        assert S.shape[1] == M
        assert S_new.shape[1] == M
        assert ray_voxel_indices.shape[1:] == (M, 3)
        assert len(ray_voxel_count.shape) == 1
        assert len(ray_voxel_count) == len(S) == len(ray_voxel_indices)
        assert len(ray_voxel_count) == len(ray_to_occupancy_messages_pon)
        assert S.shape[1] == ray_to_occupancy_messages_pon.shape[1]
        assert ray_to_occupancy_accumulated_pon.shape == tuple(grid_shape)
        assert np.float32 == S.dtype
        assert np.float32 == S_new.dtype
        assert np.float32 == ray_to_occupancy_messages_pon.dtype
        assert np.int32 == ray_voxel_indices.dtype
        assert np.int32 == ray_voxel_count.dtype
        assert acc_dtype == ray_to_occupancy_accumulated_pon.dtype
        _lib.call("rn_depth_estimate" + suffix, params, ptr(S), ptr(ray_voxel_indices), ptr(ray_voxel_count),
                  ptr(ray_to_occupancy_accumulated_pon), ptr(ray_to_occupancy_messages_pon), ptr(S_new),
                  len(S), current_stream_ptr())
        return S_new

    return bp, de


def belief_propagation(S, ray_voxel_indices, ray_voxel_count, ray_to_occupancy_messages_pon, grid_shape,
                       gamma=0.05, bp_iterations=3, batch_size=50000, parity=False):
    """mrf_cuda.py:127-197.  Host arrays in, (accumulated ndarray f32[grid], messages ndarray) out.
    parity=True (extension): float64 accumulators like mrf_np.py under NumPy >= 2; returns float64."""
    N, M = S.shape
    ray_to_occupancy_messages_pon.fill(0)
    prior = np.log(gamma) - np.log(1 - gamma)
    dt = np.float64 if parity else np.float32
    acc = to_gpu(np.full(tuple(grid_shape), prior, dtype=dt))
    acc_out = to_gpu(np.full(tuple(grid_shape), prior, dtype=dt))
    bp, _ = batch_ray_belief_propagation(M, grid_shape, parity=parity)
    for it in range(bp_iterations):
        for i in range(0, N, batch_size):
            _, msgs = bp(S[i:i + batch_size], ray_voxel_indices[i:i + batch_size],
                         ray_voxel_count[i:i + batch_size], acc,
                         ray_to_occupancy_messages_pon[i:i + batch_size], acc_out)
            ray_to_occupancy_messages_pon[i:i + batch_size] = msgs.get()
        acc_out, acc = acc, acc_out
        acc_out.fill(prior)
    return acc.get(), ray_to_occupancy_messages_pon


def compute_depth_distribution(S, ray_voxel_indices, ray_voxel_count, ray_to_occupancy_messages_pon,
                               ray_to_occupancy_accumulated_pon, S_new, grid_shape, batch_size=50000, parity=False):
    """mrf_cuda.py:200-251."""
    N, M = S.shape
    S_new.fill(0)
    _, de = batch_ray_belief_propagation(M, grid_shape, parity=parity)
    acc = to_gpu(np.ascontiguousarray(ray_to_occupancy_accumulated_pon, dtype=np.float64 if parity else np.float32))
    for i in range(0, N, batch_size):
        s = de(S[i:i + batch_size], ray_voxel_indices[i:i + batch_size], ray_voxel_count[i:i + batch_size],
               acc, ray_to_occupancy_messages_pon[i:i + batch_size], S_new[i:i + batch_size])
        S_new[i:i + batch_size] = s.get()
    return S_new


def compute_occupancy_probabilities(ray_to_occupancy_accumulated_pon, gamma=0.031):
    """mrf_np.py:206-240 on the device: sigmoid of the accumulated log-odds (in the array's own
    precision, like the numpy code: float64 accumulators give a float64 sigmoid)."""
    f64 = np.asarray(ray_to_occupancy_accumulated_pon).dtype == np.float64
    acc = to_gpu(np.ascontiguousarray(ray_to_occupancy_accumulated_pon, dtype=np.float64 if f64 else np.float32))
    out = to_gpu(np.zeros(acc.shape, dtype=np.float32))
    _lib.call("rn_occupancy_f64" if f64 else "rn_occupancy", ptr(acc), ptr(out), acc.size, current_stream_ptr())
    return out.get()
