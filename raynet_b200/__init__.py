"""raynet_b200 -- B200-native (sm_100a) implementation of RayNet's volumetric-inference hot path.

Per-ray voxel-grid DDA traversal, plane-sweep multi-view similarity mapped onto the
traversed voxels, and ray-potential sum-product belief propagation, as hand-written CUDA
behind a C-ABI shared library (include/raynet_b200.h), exposed through the reference's own
`cuda_implementations` / `mrf` / `ray_marching` / `planes_voxels_mapping` / `forward_pass`
call surface.  There is no CPU fallback: importing the compute entry points without the
built library raises.
"""
__version__ = "0.1.0"
