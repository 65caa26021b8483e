"""Drop-in for raynet/mrf/bp_inference.py: the backend strategy objects.

Only the CUDA backend exists here (there is deliberately no CPU path in the product);
asking for "numpy" or "tf" raises NotImplementedError.
"""
import numpy as np

from .mrf_cuda import belief_propagation as cuda_bp
from .mrf_cuda import compute_depth_distribution as cuda_compute_depth_distribution


class BPInference(object):
    """bp_inference.py:14-147."""

    def __init__(self, generation_params, bp_iterations=3, gamma_prior=0.05):
        self._generation_params = generation_params
        self.bp_iterations = bp_iterations
        self.gamma_prior = gamma_prior

    def update_bp_messages(self, S, ray_voxel_indices, ray_voxel_count, ray_to_occupancy_pon=None):
        raise NotImplementedError

    def estimate_depth_probabilities_from_messages(self, S, ray_voxel_indices, ray_voxel_count,
                                                   ray_to_occupancy_accumulated_pon, ray_to_occupancy_pon,
                                                   S_new):
        raise NotImplementedError

    def mrf_inference(self, S, ray_voxel_indices, ray_voxel_count, ray_to_occupancy_pon=None, S_new=None):
        ray_to_occupancy_accumulated_pon, ray_to_occupancy_pon = self.update_bp_messages(
            S, ray_voxel_indices, ray_voxel_count, ray_to_occupancy_pon)
        S_new = self.estimate_depth_probabilities_from_messages(
            S, ray_voxel_indices, ray_voxel_count, ray_to_occupancy_accumulated_pon, ray_to_occupancy_pon,
            S_new)
        return ray_to_occupancy_accumulated_pon, ray_to_occupancy_pon, S_new


class CUDABPInference(BPInference):
    """bp_inference.py:340-409."""

    def __init__(self, generation_params, batch_size=1, bp_iterations=3, gamma_prior=0.05):
        super(CUDABPInference, self).__init__(generation_params, bp_iterations, gamma_prior)
        self.batch_size = batch_size

    def update_bp_messages(self, S, ray_voxel_indices, ray_voxel_count, ray_to_occupancy_pon):
        assert S.shape[0] == ray_voxel_indices.shape[0]
        assert S.shape[0] == ray_voxel_count.shape[0]
        assert S.shape[0] == ray_to_occupancy_pon.shape[0]
        assert S.shape[1] == ray_voxel_indices.shape[1]
        assert S.shape[1] == ray_to_occupancy_pon.shape[1]
        assert len(ray_voxel_count.shape) == 1
        assert np.int32 == ray_voxel_indices.dtype
        assert np.int32 == ray_voxel_count.dtype
        assert np.float32 == S.dtype
        assert np.float32 == ray_to_occupancy_pon.dtype
        return cuda_bp(S, ray_voxel_indices, ray_voxel_count, ray_to_occupancy_pon,
                       self._generation_params.grid_shape, gamma=self.gamma_prior,
                       bp_iterations=self.bp_iterations, batch_size=self.batch_size)

    def estimate_depth_probabilities_from_messages(self, S, ray_voxel_indices, ray_voxel_count,
                                                   ray_to_occupancy_accumulated_pon, ray_to_occupancy_pon,
                                                   S_new):
        return cuda_compute_depth_distribution(S, ray_voxel_indices, ray_voxel_count, ray_to_occupancy_pon,
                                               ray_to_occupancy_accumulated_pon, S_new,
                                               self._generation_params.grid_shape, self.batch_size)


def get_bp_backend(name, generation_params, **kwargs):
    """bp_inference.py:412-439."""
    bp_iterations = kwargs["bp_iterations"] if "bp_iterations" in kwargs.keys() else 3
    if name == "cuda":
        if kwargs and "batch_size" in kwargs.keys():
            return CUDABPInference(generation_params, kwargs["batch_size"], bp_iterations=bp_iterations)
        raise ValueError("Missing argument for CUDA backend")
    if name in ("numpy", "tf"):
        raise NotImplementedError(
            "raynet_b200 ships the CUDA backend only; the %r backend of the reference is not reproduced" % name)
    return None
