"""Drop-in for the backend selector of raynet/ray_marching/ray_marching.py:84-90."""
from .ray_tracing_cuda import perform_ray_marching as perform_ray_marching_cuda


def get_voxel_traversal_backend(name):
    if name == "cuda":
        return perform_ray_marching_cuda
    elif name == "cython":
        raise NotImplementedError("raynet_b200 ships the CUDA backend only (no CPU path in the product)")
    else:
        raise NotImplementedError()
