"""RayPotentialEngine -- the resident B200 pipeline.

Replaces the host loop of RayNetForwardPass.forward_pass (raynet/forward_pass.py:593-748),
which re-runs CNN + similarity + DDA + mapping on every BP sweep and bounces the messages
through host memory per batch.  Here the front end runs ONCE per reference image and
leaves per-ray state in HBM (step codes, clip_and_renorm'ed voxel distribution, messages);
each BP sweep is a handful of kernel launches (one per ray-length class) over all rays of
this rank; the two accumulator grids are double-buffered on the device in a bricked layout.
The algorithm is the one of mrf_np.belief_propagation (mrf_np.py:243-330): synchronous
sweeps, acc_prev <- acc_new, acc_new <- prior.

Multi-GPU (one process per GPU, torch.distributed / NCCL): rays are sharded, every rank
accumulates a partial grid and the partials are summed with one all-reduce per sweep
(valid because sweeps are synchronous/Jacobi).  Rank 0 seeds its partial
with the prior so that the all-reduce result already is prior + sum of messages.
"""
import ctypes

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
        (raynet_fp.py:10-41).  Rows of the per-ray state are padded to rn_row_stride(M) floats.
        max_rays: capacity of the per-ray state on this rank."""
        M = int(M)
        self.M, self.D, self.V, self.F, self.H, self.W, self.padding = M, D, n_views, F, H, W, padding
        self.grid_shape = tuple(int(g) for g in np.asarray(grid_shape).ravel())
        self.bbox = np.asarray(bbox, dtype=np.float32).ravel()
        self.gamma = float(gamma)
        # float32 prior exactly as np.ones(f32) * (log g - log(1-g)) cast to f32 (mrf_np.py:285-292)
        self.prior = float(np.float32(np.log(self.gamma) - np.log(1 - self.gamma)))
        self.params = _lib.make_params(M, D, n_views, F, H, W, padding, self.bbox, self.grid_shape)
        self.dev = device()
        self.code_stride = _lib.code_stride(M)
        self.R = _lib.row_stride(M)             # floats per s_hat / msgs row
        self.n_classes = _lib.num_classes()
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
        self.G = int(np.prod(self.grid_shape))
        self.GB = _lib.brick_elems(self.params)     # floats of a bricked accumulator (padding included)
        kw = dict(device=self.dev)
        n = self.capacity
        self.hdr = torch.zeros((n, 2), dtype=torch.int32, **kw)
        self.codes = torch.zeros((n, self.code_stride), dtype=torch.uint8, **kw)
        self.count = torch.zeros((n,), dtype=torch.int32, **kw)
        self.lin = torch.empty((n, self.R), dtype=torch.int32, **kw)
        self.s_hat = torch.empty((n, self.R), dtype=torch.float32, **kw)
        self.msgs = torch.empty((n, self.R), dtype=torch.float32, **kw)
        self.order = torch.zeros((n,), dtype=torch.int32, **kw)
        self.acc_prev = torch.full((self.GB,), self.prior, dtype=torch.float32, **kw)
        self._acc_uniform = True         # acc_prev holds the prior everywhere (until a sweep or set_accumulator)
        self.acc_new = torch.empty((self.GB,), dtype=torch.float32, **kw)
        self.axes = torch.zeros((sum(self.grid_shape),), dtype=torch.float32, **kw)
        self._side = None                # side stream + pinned buffer for the class-size read-back
        self._sizes_host = None
        self._pending_sizes = None
        self._binned = False
        self.starts = self.ends = None   # float32 [capacity, 3], allocated by the first trace_image()
        self._class_scratch = torch.zeros((2 * self.n_classes,), dtype=torch.int64, **kw)
        self._class_offsets = None  # host int64 [n_classes + 1] (ctypes array) once rays are binned
        self.class_sizes = None
        self._centres = None
        self._seg_starts = None
        self._axes_set = False
        self.iterations_done = 0
        self.sweep_events = None    # bench.py: list of (start, end) CUDA events around each sweep

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
        """Forget the rays; messages count as 0 (the first sweep does not read them) and the
        accumulator is back at the prior (mrf_np.py:275-292)."""
        _lib.call("rn_fill_f32", _ptr(self.acc_prev), self.prior, self.GB, current_stream_ptr())
        self._acc_uniform = True
        self.launches += 1
        self.iterations_done = 0
        self.n_rays = 0
        self.segments = []
        self._class_offsets = None
        self._binned = False
        self._pending_sizes = None
        self.class_sizes = None
        self._centres = None
        self._seg_starts = None

    # ------------------------------------------------------------------ front end
    def add_image(self, ray_idxs, features, P, P_inv, centre, view_ids=None, n_feature_slots=None,
                  keep_start_end=False):
        """Front end for the rays `ray_idxs` (int32 device tensor, column-major pixel ids) of one
        reference image.  features: CUDA f32 tensor [V or slots, H+p+1, W+p+1, F]; P [V,3,4],
        P_inv [4,3], centre [4] CUDA f32 tensors.  Appends the rays to the resident state."""
        k = self.trace_image(ray_idxs, P_inv, centre)
        self.score_image(k, features, P, view_ids=view_ids, n_feature_slots=n_feature_slots)
        if keep_start_end:
            start, n, _ = self.segments[k]
            return self.starts[start:start + n], self.ends[start:start + n]
        return None

    def trace_image(self, ray_idxs, P_inv, centre):
        """First half of the front end (needs no feature maps): sample_in_bbox + DDA for the rays
        of one reference image -> step codes, counts, ray start / end.  Returns the image's index."""
        assert self._axes_set, "call set_voxel_grid() first"
        n = int(ray_idxs.shape[0])
        start = self.n_rays
        if start + n > self.capacity:
            raise AssertionError("engine capacity exceeded: %d + %d > %d" % (start, n, self.capacity))
        sl = slice(start, start + n)
        if self.starts is None:
            self.starts = torch.empty((self.capacity, 3), dtype=torch.float32, device=self.dev)
            self.ends = torch.empty((self.capacity, 3), dtype=torch.float32, device=self.dev)
        _lib.call("rn_engine_trace", self.params, _ptr(ray_idxs), _ptr(P_inv), _ptr(centre), _ptr(self.starts[sl]),
                  _ptr(self.ends[sl]), _ptr(self.hdr[sl]), _ptr(self.codes[sl]), _ptr(self.count[sl]), n,
                  current_stream_ptr())
        self.launches += 1
        self.segments.append((start, n, centre))
        self.n_rays = start + n
        self._class_offsets = None
        self._binned = False
        self._pending_sizes = None
        return len(self.segments) - 1

    def score_image(self, k, features, P, view_ids=None, n_feature_slots=None):
        """Second half of the front end for image k of trace_image(): plane-sweep similarity +
        plane->voxel mapping -> s_hat, lin rows."""
        start, n, _ = self.segments[k]
        sl = slice(start, start + n)
        slots = int(n_feature_slots if n_feature_slots is not None else features.shape[0])
        _lib.call("rn_engine_similarity", self.params, _ptr(features), _ptr(view_ids) if view_ids is not None else None,
                  slots, _ptr(P), _ptr(self.axes), _ptr(self.starts[sl]), _ptr(self.ends[sl]), _ptr(self.hdr[sl]),
                  _ptr(self.codes[sl]), _ptr(self.count[sl]), _ptr(self.s_hat[sl]), _ptr(self.lin[sl]), n,
                  current_stream_ptr())
        self.launches += 2

    def finalize_frontend(self):
        """Bin the rays by length class (one small device->host read of the class sizes, so
        every class is launched with exactly the shared memory its rays need) and stage the
        per-image camera centres for the depth pass."""
        st = current_stream_ptr()
        sizes = set(n for (_, n, _) in self.segments)
        seg_len = sizes.pop() if len(sizes) == 1 else 0
        _lib.call("rn_engine_bin_rays", self.params, _ptr(self.count), self.n_rays, int(seg_len), _ptr(self.order),
                  _ptr(self._class_scratch), st)
        self.launches += 2
        cen = torch.stack([torch.cat([c.reshape(-1)[:3], c.new_ones(1)]) for (_, _, c) in self.segments]).contiguous()
        self._centres = cen
        self._seg_starts = torch.tensor([s for (s, _, _) in self.segments] + [self.n_rays], dtype=torch.int64,
                                        device=self.dev)
        # the one device->host read: on a side stream into pinned memory, so that kernels the caller
        # queues on its own stream meanwhile (the similarity of forward_pass / bench) are not waited for
        main = torch.cuda.current_stream(self.dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
            self._sizes_host = torch.empty((self.n_classes,), dtype=torch.int64).pin_memory()
        binned = torch.cuda.Event()
        binned.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(binned)
            self._sizes_host.copy_(self._class_scratch[:self.n_classes], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._side)
        self._pending_sizes = done
        self._binned = True

    def _resolve_classes(self):
        """Wait for the class sizes read back by finalize_frontend() (normally long complete: the
        caller has queued the similarity kernels in the meantime) and turn them into launch offsets."""
        if not self._binned:
            self.finalize_frontend()
        if self._pending_sizes is None:
            return self.max_count
        self._pending_sizes.synchronize()
        self._pending_sizes = None
        sizes = self._sizes_host.numpy().astype(np.int64)
        self.class_sizes = sizes
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        self._class_offsets = (ctypes.c_int64 * (self.n_classes + 1))(*[int(v) for v in off])
        nz = np.nonzero(sizes[1:])[0]
        self.max_count = int(min(self.M, 128 * (int(nz[-1]) + 1))) if len(nz) else 1
        return self.max_count

    # ------------------------------------------------------------------ BP
    def bp_iteration(self):
        if self._class_offsets is None:
            self._resolve_classes()
        st = current_stream_ptr()
        _lib.call("rn_fill_f32", _ptr(self.acc_new), sharding.seed_value(self.rank, self.prior), self.GB, st)
        if self.sweep_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _lib.call("rn_engine_bp_iteration", self.params, _ptr(self.lin), _ptr(self.count),
                  _ptr(self.s_hat), _ptr(self.msgs), _ptr(self.acc_prev), _ptr(self.acc_new), _ptr(self.order),
                  self._class_offsets, (2 if self._acc_uniform else 1) if self.iterations_done == 0 else 0,
                  int(self.max_count), self.n_rays, st)
        self._acc_uniform = False
        if self.sweep_events is not None:
            ev[1].record()
            self.sweep_events.append(ev)
        self.launches += 1 + int(np.count_nonzero(self.class_sizes[1:]))
        if self.world > 1:
            sharding.allreduce_accumulator(self.acc_new, self.pg)
        self.acc_prev, self.acc_new = self.acc_new, self.acc_prev
        self.iterations_done += 1

    def run_bp(self, iterations):
        for _ in range(int(iterations)):
            self.bp_iteration()
        return self.acc_prev

    # ------------------------------------------------------------------ state access (row-major views)
    def accumulator(self):
        """acc_prev as the reference's row-major float32 [Gx, Gy, Gz] grid."""
        out = torch.empty(self.grid_shape, dtype=torch.float32, device=self.dev)
        _lib.call("rn_bricks_to_grid", self.params, _ptr(self.acc_prev), _ptr(out), 0, current_stream_ptr())
        self.launches += 1
        return out

    def set_accumulator(self, grid):
        """Load a row-major [Gx, Gy, Gz] grid (CUDA tensor or numpy) as acc_prev."""
        if isinstance(grid, np.ndarray):
            grid = torch.from_numpy(np.ascontiguousarray(grid, dtype=np.float32)).to(self.dev)
        grid = grid.reshape(self.grid_shape).contiguous()
        _lib.call("rn_grid_to_bricks", self.params, _ptr(grid), _ptr(self.acc_prev), self.prior, current_stream_ptr())
        self._acc_uniform = False
        self.launches += 1

    def set_messages(self, msgs):
        """Load messages [n_rays, <= M] and mark the state as past the first sweep."""
        if isinstance(msgs, np.ndarray):
            msgs = torch.from_numpy(np.ascontiguousarray(msgs, dtype=np.float32)).to(self.dev)
        self.msgs[:msgs.shape[0], :msgs.shape[1]].copy_(msgs)
        self.iterations_done = max(self.iterations_done, 1)

    def messages(self):
        if self.iterations_done == 0:
            return torch.zeros((self.n_rays, self.M), dtype=torch.float32, device=self.dev)
        return self.msgs[:self.n_rays, :self.M]

    # ------------------------------------------------------------------ outputs
    def depth(self, depth_out=None, S_new=None):
        """Depth per ray (flat, ray order of add_image calls), all images in one launch."""
        if self._class_offsets is None:
            self._resolve_classes()
        if depth_out is None:
            depth_out = torch.empty((self.n_rays,), dtype=torch.float32, device=self.dev)
        if self.iterations_done == 0:
            self.msgs[:self.n_rays].zero_()
        _lib.call("rn_engine_depth", self.params, _ptr(self.lin), _ptr(self.count), _ptr(self.s_hat),
                  _ptr(self.msgs), _ptr(self.acc_prev), _ptr(self.axes), _ptr(self._centres), _ptr(self._seg_starts),
                  len(self.segments), _ptr(depth_out), _ptr(S_new), self.n_rays, current_stream_ptr())
        self.launches += 1
        return depth_out

    def depth_distribution(self):
        """S_new [n_rays, M] (compute_depth_distribution, mrf_np.py:333-385) -- parity tests."""
        S_new = torch.empty((self.n_rays, self.R), dtype=torch.float32, device=self.dev)
        self.depth(S_new=S_new)
        return S_new[:, :self.M]

    def occupancy(self):
        """sigmoid(acc_prev) as a row-major [Gx, Gy, Gz] grid (mrf_np.py:206-240)."""
        out = torch.empty(self.grid_shape, dtype=torch.float32, device=self.dev)
        _lib.call("rn_bricks_to_grid", self.params, _ptr(self.acc_prev), _ptr(out), 1, current_stream_ptr())
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
