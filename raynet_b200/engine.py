"""RayPotentialEngine -- the resident B200 pipeline.

Replaces the host loop of RayNetForwardPass.forward_pass (raynet/forward_pass.py:593-748),
which re-runs CNN + similarity + DDA + mapping on every BP sweep and bounces the messages
through host memory per batch.  Here the front end runs ONCE per reference image and
leaves per-ray state in HBM (step codes, clip_and_renorm'ed voxel distribution, messages);
each BP sweep is one kernel launch over all rays of this rank; the two accumulator grids
are double-buffered on the device.  The algorithm is the one of mrf_np.belief_propagation
(mrf_np.py:243-330): synchronous sweeps, acc_prev <- acc_new, acc_new <- prior.

Multi-GPU (one process per GPU, torch.distributed / NCCL): rays are sharded, every rank
accumulates a partial grid and the partials are summed with one all-reduce per sweep
(valid because sweeps are synchronous/Jacobi).  Rank 0 seeds its partial
with the prior so that the all-reduce result already is prior + sum of messages.
"""
import numpy as np
import torch

from . import _lib, sharding
from .cuda_implementations.utils import current_stream_ptr, device


def _ptr(t):
    if t is None:
        return None
    assert t.is_contiguous()
    return t.data_ptr()


class RayPotentialEngine(object):
    def __init__(self, M, D, n_views, F, H, W, padding, bbox, grid_shape, gamma=0.05, max_rays=0,
                 process_group=None, use_distributed=None):
        """M, D, n_views, F, H, W, padding, bbox, grid_shape: as perform_raynet_fp
        (raynet_fp.py:10-41).  max_rays: capacity of the per-ray state on this rank."""
        if M % 4 != 0:
            raise AssertionError("resident layout needs max_voxels to be a multiple of 4")
        self.M, self.D, self.V, self.F, self.H, self.W, self.padding = M, D, n_views, F, H, W, padding
        self.grid_shape = tuple(int(g) for g in np.asarray(grid_shape).ravel())
        self.bbox = np.asarray(bbox, dtype=np.float32).ravel()
        self.gamma = float(gamma)
        # float32 prior exactly as np.ones(f32) * (log g - log(1-g)) cast to f32 (mrf_np.py:285-292)
        self.prior = float(np.float32(np.log(self.gamma) - np.log(1 - self.gamma)))
        self.params = _lib.make_params(M, D, n_views, F, H, W, padding, self.bbox, self.grid_shape)
        self.dev = device()
        self.code_stride = _lib.code_stride(M)
        self.capacity = int(max_rays)
        self.n_rays = 0
        self.max_count = M
        self.segments = []          # (start, n, centre_tensor) per reference image
        self.pg = process_group
        if use_distributed is None:
            use_distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.distributed = bool(use_distributed)
        self.rank = torch.distributed.get_rank(self.pg) if self.distributed else 0
        self.world = torch.distributed.get_world_size(self.pg) if self.distributed else 1
        self.launches = 0           # kernels launched by this engine (bench.py's gpu_launches)
        G = int(np.prod(self.grid_shape))
        self.G = G
        kw = dict(device=self.dev)
        n = self.capacity
        self.hdr = torch.zeros((n, 2), dtype=torch.int32, **kw)
        self.codes = torch.zeros((n, self.code_stride), dtype=torch.uint8, **kw)
        self.count = torch.zeros((n,), dtype=torch.int32, **kw)
        self.s_hat = torch.zeros((n, M), dtype=torch.float32, **kw)
        self.msgs = torch.zeros((n, M), dtype=torch.float32, **kw)
        self.acc_prev = torch.full(self.grid_shape, self.prior, dtype=torch.float32, **kw)
        self.acc_new = torch.empty(self.grid_shape, dtype=torch.float32, **kw)
        self.axes = torch.zeros((sum(self.grid_shape),), dtype=torch.float32, **kw)
        self._max_count_dev = torch.zeros((1,), dtype=torch.int32, **kw)
        self._axes_set = False
        self.iterations_done = 0
        self.sweep_events = None    # bench.py: list of (start, end) CUDA events around each sweep kernel

    # ------------------------------------------------------------------ setup
    def set_voxel_grid(self, voxel_grid):
        """voxel_grid: the reference's table, (3, Gx, Gy, Gz) as Scene.voxel_grid returns it or
        (Gx, Gy, Gz, 3) as it is handed to the kernels (forward_pass.py:571-576); numpy or a
        CUDA tensor.  Only its three axis slices are kept (the table is separable)."""
        if isinstance(voxel_grid, np.ndarray):
            vg = voxel_grid
            if vg.shape[0] == 3 and vg.ndim == 4 and vg.shape[1:] == self.grid_shape:
                ax = np.concatenate([vg[0, :, 0, 0], vg[1, 0, :, 0], vg[2, 0, 0, :]]).astype(np.float32)
            else:
                vg = vg.reshape(self.grid_shape + (3,))
                ax = np.concatenate([vg[:, 0, 0, 0], vg[0, :, 0, 1], vg[0, 0, :, 2]]).astype(np.float32)
            self.axes.copy_(torch.from_numpy(ax))
        else:
            t = voxel_grid.reshape(self.grid_shape + (3,)).contiguous()
            _lib.call("rn_axis_centres", self.params, _ptr(t), _ptr(self.axes), current_stream_ptr())
            self.launches += 1
        self._axes_set = True

    def reset(self):
        """Messages to 0, accumulator to the prior (mrf_np.py:275-292)."""
        self.msgs.zero_()
        self.acc_prev.fill_(self.prior)
        self.iterations_done = 0
        self.n_rays = 0
        self.segments = []

    # ------------------------------------------------------------------ front end
    def add_image(self, ray_idxs, features, P, P_inv, centre, view_ids=None, n_feature_slots=None,
                  keep_start_end=False):
        """Front end for the rays `ray_idxs` (int32 device tensor, column-major pixel ids) of one
        reference image.  features: CUDA f32 tensor [V or slots, H+p+1, W+p+1, F]; P [V,3,4],
        P_inv [4,3], centre [4] CUDA f32 tensors.  Appends the rays to the resident state."""
        assert self._axes_set, "call set_voxel_grid() first"
        n = int(ray_idxs.shape[0])
        start = self.n_rays
        if start + n > self.capacity:
            raise AssertionError("engine capacity exceeded: %d + %d > %d" % (start, n, self.capacity))
        sl = slice(start, start + n)
        starts = ends = None
        if keep_start_end:
            starts = torch.empty((n, 3), dtype=torch.float32, device=self.dev)
            ends = torch.empty((n, 3), dtype=torch.float32, device=self.dev)
        slots = int(n_feature_slots if n_feature_slots is not None else features.shape[0])
        _lib.call("rn_engine_frontend", self.params, _ptr(ray_idxs), _ptr(features),
                  _ptr(view_ids) if view_ids is not None else None, slots, _ptr(P), _ptr(P_inv), _ptr(centre),
                  _ptr(self.axes), _ptr(starts), _ptr(ends), _ptr(self.hdr[sl]), _ptr(self.codes[sl]),
                  _ptr(self.count[sl]), _ptr(self.s_hat[sl]), n, current_stream_ptr())
        self.launches += 2
        self.segments.append((start, n, centre))
        self.n_rays = start + n
        return (starts, ends) if keep_start_end else None

    def finalize_frontend(self):
        """One device->host read of the longest ray so the sweep kernels are instantiated for
        the actual ray length instead of the capacity M."""
        _lib.call("rn_max_count", _ptr(self.count), self.n_rays, _ptr(self._max_count_dev), current_stream_ptr())
        self.launches += 1
        self.max_count = max(1, int(self._max_count_dev.item()))
        return self.max_count

    # ------------------------------------------------------------------ BP
    def bp_iteration(self):
        st = current_stream_ptr()
        _lib.call("rn_fill_f32", _ptr(self.acc_new), sharding.seed_value(self.rank, self.prior), self.G, st)
        if self.sweep_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _lib.call("rn_engine_bp_iteration", self.params, _ptr(self.hdr), _ptr(self.codes), _ptr(self.count),
                  _ptr(self.s_hat), _ptr(self.msgs), _ptr(self.acc_prev), _ptr(self.acc_new),
                  int(self.max_count), self.n_rays, st)
        if self.sweep_events is not None:
            ev[1].record()
            self.sweep_events.append(ev)
        self.launches += 2
        if self.world > 1:
            sharding.allreduce_accumulator(self.acc_new, self.pg)
        self.acc_prev, self.acc_new = self.acc_new, self.acc_prev
        self.iterations_done += 1

    def run_bp(self, iterations):
        for _ in range(int(iterations)):
            self.bp_iteration()
        return self.acc_prev

    # ------------------------------------------------------------------ outputs
    def depth(self, depth_out=None):
        """Depth per ray (flat, ray order of add_image calls)."""
        if depth_out is None:
            depth_out = torch.empty((self.n_rays,), dtype=torch.float32, device=self.dev)
        st = current_stream_ptr()
        for (start, n, centre) in self.segments:
            sl = slice(start, start + n)
            _lib.call("rn_engine_depth", self.params, _ptr(self.hdr[sl]), _ptr(self.codes[sl]), _ptr(self.count[sl]),
                      _ptr(self.s_hat[sl]), _ptr(self.msgs[sl]), _ptr(self.acc_prev), _ptr(self.axes), _ptr(centre),
                      _ptr(depth_out[sl]), int(self.max_count), n, st)
            self.launches += 1
        return depth_out

    def occupancy(self):
        out = torch.empty_like(self.acc_prev)
        _lib.call("rn_occupancy", _ptr(self.acc_prev), _ptr(out), self.G, current_stream_ptr())
        self.launches += 1
        return out

    def voxel_indices(self, start=0, n=None):
        """Dense int32 [n, M, 3] lists (reference layout) expanded from the step codes."""
        n = self.n_rays - start if n is None else n
        sl = slice(start, start + n)
        out = torch.empty((n, self.M, 3), dtype=torch.int32, device=self.dev)
        _lib.call("rn_engine_expand_indices", self.params, _ptr(self.hdr[sl]), _ptr(self.codes[sl]),
                  _ptr(self.count[sl]), _ptr(out), n, current_stream_ptr())
        self.launches += 1
        return out
