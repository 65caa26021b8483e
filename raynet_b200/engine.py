"""RayPotentialEngine -- the resident B200 pipeline.

Replaces the host loop of RayNetForwardPass.forward_pass (raynet/forward_pass.py:593-748),
which re-runs CNN + similarity + DDA + mapping on every BP sweep and bounces the messages
through host memory per batch.  Here the front end runs ONCE per reference image and
leaves per-ray state in HBM (step codes, clip_and_renorm'ed voxel distribution, messages);
each BP sweep is a handful of kernel launches (one per ray-length class) over all rays of
this rank; the two accumulator grids are double-buffered on the device in a bricked layout.
The algorithm is the one of mrf_np.belief_propagation (mrf_np.py:243-330): synchronous
sweeps, acc_prev <- acc_new, acc_new <- prior.

Segments.  Rays are added in SEGMENTS: the rays of one reference image, or any part of them
(a rank of a multi-GPU job owns a contiguous block of the (image, column-major pixel) ray
enumeration, so its first and last segments are partial images).

Bounded memory.  The reference processes `rays_batch` rays at a time and spills the messages
to disk (forward_pass.py:588,602-611).  Here the per-ray state is 12 R + 0.25 M + 44 bytes per
ray (R = M rounded up to 128).  When `memory_budget` does not hold that for all rays, the
leading segments stay fully RESIDENT and the others are STREAMED: only their step codes and
messages stay in HBM (4 R + 0.25 M + 44 bytes per ray) and their `lin` / `s_hat` rows are
recomputed into a window (one segment long) on every sweep -- what the reference does for every
ray on every sweep.  If even that does not fit, the constructor raises MemoryError with the
numbers instead of running into a CUDA out-of-memory error.

Parity mode (`parity=True`): float64 accumulators and a float64
occupancy-to-ray chain, the arithmetic of mrf_np.py as it executes under NumPy >= 2
(csrc/rn_parity.cuh).  The checking mode behind the end-to-end 1e-5 gate; ~4x slower sweeps.

Multi-GPU (one process per GPU, torch.distributed / NCCL): rays are sharded, every rank
accumulates a partial grid and the partials are summed with one all-reduce per sweep
(valid because sweeps are synchronous/Jacobi).  Rank 0 seeds its partial
with the prior so that the all-reduce result already is prior + sum of messages.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib, sharding
from .cuda_implementations.utils import current_stream_ptr, device


def _ptr(t):
    if t is None:
        return None
    assert t.is_contiguous()
    return t.data_ptr()


class _Group(object):
    """A set of consecutive segments swept together: rays [start, start + n) of the engine."""

    def __init__(self, start, n, segs, resident):
        self.start, self.n, self.segs, self.resident = start, n, segs, resident
        self.class_offsets = None      # ctypes int64 [n_classes + 1]
        self.class_sizes = None
        self.centres = None            # float32 [n_seg, 4]
        self.seg_starts = None         # int64 [n_seg + 1], relative to the group


class RayPotentialEngine(object):
    def __init__(self, M, D, n_views, F, H, W, padding, bbox, grid_shape, gamma=0.05, max_rays=0,
                 process_group=None, use_distributed=None, parity=False, memory_budget=None,
                 max_segment_rays=None, collective="auto", fuse_first_sweep=False):
        """M, D, n_views, F, H, W, padding, bbox, grid_shape: as perform_raynet_fp
        (raynet_fp.py:10-41).  Rows of the per-ray state are padded to rn_row_stride(M) floats.
        max_rays: capacity of the per-ray state on this rank.  memory_budget: bytes of HBM the
        per-ray state may take (default: 85 % of the free memory at construction).
        max_segment_rays: upper bound on the rays of one segment (default H * W).
        collective: how the per-rank partial accumulators are summed after a sweep when world > 1 --
        "peer": this library's fused exchange kernel over NVLink peer memory (sharding.PeerExchange; with the sum and
        the broadcast done inside the NVSwitch where the fabric offers multicast, "peer_p2p" keeps it on peer loads),
        "nccl": torch.distributed all_reduce, "auto": peer when the GPUs can map each other's memory
        (float32 accumulators only), else nccl; self.collective says which one runs.
        fuse_first_sweep: score_image() only computes the plane distributions; the FIRST sweep after a reset
        builds the voxel-space rows inside the sweep kernel (csrc/rn_first.cuh) -- for callers that go
        front end -> sweeps -> depth (forward_pass, bench); rows are materialised on demand otherwise."""
        M = int(M)
        self.M, self.D, self.V, self.F, self.H, self.W, self.padding = M, D, n_views, F, H, W, padding
        self.grid_shape = tuple(int(g) for g in np.asarray(grid_shape).ravel())
        self.bbox = np.asarray(bbox, dtype=np.float32).ravel()
        self.gamma = float(gamma)
        self.parity = bool(parity)
        self.fuse_first = bool(fuse_first_sweep) and not self.parity and int(F) == 32
        # float32 prior exactly as np.ones(f32) * (log g - log(1-g)) cast to f32 (mrf_np.py:285-292);
        # parity mode keeps the float64 value like NumPy >= 2 does
        self.prior64 = float(np.log(self.gamma) - np.log(1 - self.gamma))
        self.prior = self.prior64 if self.parity else float(np.float32(self.prior64))
        self.params = _lib.make_params(M, D, n_views, F, H, W, padding, self.bbox, self.grid_shape)
        self.dev = device()
        self.code_stride = _lib.code_stride(M)
        self.R = _lib.row_stride(M)             # floats per s_hat / msgs row
        self.n_classes = _lib.num_classes()
        self.capacity = int(max_rays)
        self.n_rays = 0
        self.max_count = M
        self.segments = []          # (start, n, centre_tensor) per segment
        self.pg = process_group
        if use_distributed is None:
            use_distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.distributed = bool(use_distributed)
        self.rank = torch.distributed.get_rank(self.pg) if self.distributed else 0
        self.world = torch.distributed.get_world_size(self.pg) if self.distributed else 1
        self.launches = 0           # kernels launched by this engine (bench.py's gpu_launches)
        self.G = int(np.prod(self.grid_shape))
        self.GB = _lib.brick_elems(self.params)     # elements of a bricked accumulator (padding included)
        kw = dict(device=self.dev)
        n = self.capacity
        self.max_segment_rays = int(min(n, max_segment_rays if max_segment_rays else H * W)) if n else 0
        self._plan_memory(memory_budget)
        nr = self.resident_capacity
        self.hdr = torch.zeros((n, 2), dtype=torch.int32, **kw)
        self.codes = torch.zeros((n, self.code_stride), dtype=torch.uint8, **kw)
        self.count = torch.zeros((n,), dtype=torch.int32, **kw)
        self.msgs = torch.empty((n, self.R), dtype=torch.float32, **kw)
        self.order = torch.zeros((n,), dtype=torch.int32, **kw)
        # lin / s_hat: rows of the resident rays, followed by the window of the streamed segments
        rows = nr + (self.max_segment_rays if nr < n else 0)
        self.lin = torch.empty((rows, self.R), dtype=torch.int32, **kw)
        self.s_hat = torch.empty((rows, self.R), dtype=torch.float32, **kw)
        acc_dtype = torch.float64 if self.parity else torch.float32
        self.collective = "none"
        self._peer = None
        if self.world > 1:
            assert collective in ("auto", "peer", "peer_p2p", "nccl")
            self.collective = "nccl"
            if collective in ("auto", "peer", "peer_p2p") and not self.parity:
                try:
                    self._peer = sharding.PeerExchange(self.GB, self.dev, self.pg, multicast=(False if collective == "peer_p2p" else None))
                    self.collective = "peer_multicast" if self._peer.multicast else "peer"
                except RuntimeError as e:
                    if collective != "auto":
                        raise
                    self.collective = "nccl (%s)" % (e,)
        if self._peer is not None:
            # fixed roles: sweeps read `result` and scatter-add into `partial`; the exchange kernel refills `result`
            # (the partial is double-buffered and cleared by the exchange kernel itself: no fill between sweeps)
            self.acc_prev, self.acc_new = self._peer.result, self._peer.partial
            self.acc_prev.fill_(self.prior)
        else:
            self.acc_prev = torch.full((self.GB,), self.prior, dtype=acc_dtype, **kw)
            self.acc_new = torch.empty((self.GB,), dtype=acc_dtype, **kw)
        self._acc_uniform = True         # acc_prev holds the prior everywhere (until a sweep or set_accumulator)
        self.axes = torch.zeros((sum(self.grid_shape),), dtype=torch.float32, **kw)
        self._planes = None              # float32 [max_segment_rays, D] plane-distribution scratch of the front end
        self._planes_all = None          # fuse_first: float32 [resident rays, D] plane distributions of every resident ray
        self._unmapped = []              # fuse_first: resident segments whose lin / s_hat rows are not built yet
        self._side = None                # side stream + pinned buffer for the class-size read-back
        self._sizes_host = None
        self._pending_sizes = None
        self._binned = False
        self._classes_ready = False
        self.starts = self.ends = None   # float32 [capacity, 3], allocated by the first trace_image()
        self._class_scratch = None
        self.groups = None
        self.class_sizes = None
        self._scored = {}                # segment index -> (features, P, view_ids, slots): how to re-score a streamed segment
        self._axes_set = False
        self.iterations_done = 0
        self.sweep_events = None    # bench.py: list of (start, end) CUDA events around each sweep
        self.exchange_events = None # bench.py: the same around each exchange (wait for the slowest rank included)

    # ------------------------------------------------------------------ memory plan
    def bytes_per_ray(self, resident=True):
        light = 4 * self.R + self.code_stride + 8 + 4 + 4 + 24      # msgs, codes, hdr, count, order, starts / ends
        return light + ((8 * self.R + (4 * self.D if self.fuse_first else 0)) if resident else 0)

    def _plan_memory(self, memory_budget):
        n = self.capacity
        if memory_budget is None:
            free, _ = torch.cuda.mem_get_info(self.dev)
            memory_budget = int(0.85 * free)
        self.memory_budget = int(memory_budget)
        grids = 2 * self.GB * (8 if self.parity else 4) + self.max_segment_rays * self.D * 4
        avail = self.memory_budget - grids
        if n * self.bytes_per_ray(True) <= avail:
            self.resident_capacity = n
            return
        window = self.max_segment_rays * 8 * self.R
        light = n * self.bytes_per_ray(False)
        if light + window > avail:
            raise MemoryError(
                "raynet_b200: %d rays of up to %d voxels need %.1f GB of HBM even with every segment streamed "
                "(messages and step codes %.1f GB + window %.1f GB + grids %.1f GB) but the budget is %.1f GB; "
                "run fewer reference images per call or more GPUs" %
                (n, self.M, (light + window + grids) / 1e9, light / 1e9, window / 1e9, grids / 1e9,
                 self.memory_budget / 1e9))
        self.resident_capacity = int(max(0, (avail - light - window) // (self.bytes_per_ray(True) - self.bytes_per_ray(False))))

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

    def _fill(self, acc, value):
        _lib.call("rn_fill_f64" if self.parity else "rn_fill_f32", _ptr(acc), float(value), self.GB, current_stream_ptr())
        self.launches += 1

    def reset(self):
        """Forget the rays; messages count as 0 (the first sweep does not read them) and the
        accumulator is back at the prior (mrf_np.py:275-292)."""
        self._fill(self.acc_prev, self.prior)
        self._acc_uniform = True
        self.iterations_done = 0
        self.n_rays = 0
        self.segments = []
        self.groups = None
        self._scored = {}
        self._unmapped = []
        self._binned = False
        self._classes_ready = False
        self._pending_sizes = None
        self.class_sizes = None

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

    def is_resident(self, k):
        start, n, _ = self.segments[k]
        return start + n <= self.resident_capacity

    def trace_image(self, ray_idxs, P_inv, centre):
        """First half of the front end (needs no feature maps): sample_in_bbox + DDA for the rays
        of one segment -> step codes, counts, ray start / end.  Returns the segment's index.
        Segments can only be added before the first sweep (their messages start at 0)."""
        assert self._axes_set, "call set_voxel_grid() first"
        assert self.iterations_done == 0, "segments can only be added before the first sweep; call reset() first"
        n = int(ray_idxs.shape[0])
        start = self.n_rays
        if start + n > self.capacity:
            raise AssertionError("engine capacity exceeded: %d + %d > %d" % (start, n, self.capacity))
        if n > self.max_segment_rays:
            raise AssertionError("segment of %d rays exceeds max_segment_rays = %d" % (n, self.max_segment_rays))
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
        self.groups = None
        self._binned = False
        self._classes_ready = False
        self._pending_sizes = None
        return len(self.segments) - 1

    def _score(self, k, rows_at, features, P, view_ids, slots):
        """similarity + plane->voxel mapping of segment k into rows [rows_at, rows_at + n) of lin / s_hat."""
        start, n, _ = self.segments[k]
        if n == 0:
            return
        sl = slice(start, start + n)
        rw = slice(rows_at, rows_at + n)
        if self._planes is None:
            self._planes = torch.empty((self.max_segment_rays, self.D), dtype=torch.float32, device=self.dev)
        _lib.call("rn_engine_similarity", self.params, _ptr(features), _ptr(view_ids) if view_ids is not None else None,
                  slots, _ptr(P), _ptr(self.axes), _ptr(self.starts[sl]), _ptr(self.ends[sl]), _ptr(self.hdr[sl]),
                  _ptr(self.codes[sl]), _ptr(self.count[sl]), _ptr(self._planes), _ptr(self.s_hat[rw]),
                  _ptr(self.lin[rw]), n, current_stream_ptr())
        self.launches += 2

    def score_image(self, k, features, P, view_ids=None, n_feature_slots=None):
        """Second half of the front end for segment k of trace_image(): plane-sweep similarity +
        plane->voxel mapping -> s_hat, lin rows.  A streamed segment (see the module docstring) only
        records its inputs: it is scored into the window at every sweep.  With fuse_first_sweep only the
        plane distributions are computed here; the rows are built by the first sweep (or on demand)."""
        slots = int(n_feature_slots if n_feature_slots is not None else features.shape[0])
        self._scored[k] = (features, P, view_ids, slots)
        if not self.is_resident(k):
            return
        start, n, _ = self.segments[k]
        if self.fuse_first and n > 0:
            if self._planes_all is None:
                self._planes_all = torch.empty((self.resident_capacity, self.D), dtype=torch.float32, device=self.dev)
            sl = slice(start, start + n)
            try:
                _lib.call("rn_engine_plane_scores", self.params, _ptr(features), _ptr(view_ids) if view_ids is not None else None,
                          slots, _ptr(P), _ptr(self.starts[sl]), _ptr(self.ends[sl]), _ptr(self._planes_all[sl]), n,
                          current_stream_ptr())
                self.launches += 1
                self._unmapped.append(k)
                return
            except NotImplementedError:      # a feature volume the F = 32 kernels cannot address: unfused path
                self.fuse_first = False
                self._ensure_mapped()
        self._score(k, start, features, P, view_ids, slots)

    def _ensure_mapped(self):
        """Build the lin / s_hat rows of the resident segments that only have plane distributions so far."""
        for k in self._unmapped:
            start, n, _ = self.segments[k]
            sl = slice(start, start + n)
            _lib.call("rn_engine_map_planes", self.params, _ptr(self.axes), _ptr(self.starts[sl]), _ptr(self.ends[sl]),
                      _ptr(self.hdr[sl]), _ptr(self.codes[sl]), _ptr(self.count[sl]), _ptr(self._planes_all[sl]),
                      _ptr(self.s_hat[sl]), _ptr(self.lin[sl]), n, current_stream_ptr())
            self.launches += 1
        self._unmapped = []

    def _make_groups(self):
        groups, res = [], [k for k in range(len(self.segments)) if self.is_resident(k)]
        if res:
            a = self.segments[res[0]][0]
            b = self.segments[res[-1]][0] + self.segments[res[-1]][1]
            groups.append(_Group(a, b - a, res, True))
        for k in range(len(self.segments)):
            if not self.is_resident(k):
                groups.append(_Group(self.segments[k][0], self.segments[k][1], [k], False))
        return groups

    def finalize_frontend(self):
        """Bin the rays by length class (one small device->host read of the class sizes, so
        every class is launched with exactly the shared memory its rays need) and stage the
        per-segment camera centres for the depth pass."""
        st = current_stream_ptr()
        self.groups = self._make_groups()
        ng = max(1, len(self.groups))
        if self._class_scratch is None or self._class_scratch.shape[0] < ng:
            self._class_scratch = torch.zeros((ng, 2 * self.n_classes), dtype=torch.int64, device=self.dev)
            self._sizes_host = torch.empty((ng, 2 * self.n_classes), dtype=torch.int64).pin_memory()
        unit = 8 * self.H
        for gi, g in enumerate(self.groups):
            # the tiled enumeration works on runs of whole 8-pixel (or 64-pixel) column groups: the largest
            # common unit of the segment lengths, if there is one
            seg_len = 0
            if g.n > 0 and unit > 0:
                q = 0
                for k in g.segs:
                    q = math.gcd(q, self.segments[k][1])
                if q % unit == 0:
                    seg_len = q if q % (64 * self.H) or q == self.H * self.W else 64 * self.H
            sl = slice(g.start, g.start + g.n)
            _lib.call("rn_engine_bin_rays", self.params, _ptr(self.count[sl]), g.n, int(seg_len), _ptr(self.order[sl]),
                      _ptr(self._class_scratch[gi]), st)
            self.launches += 2
            cen = torch.stack([torch.cat([self.segments[k][2].reshape(-1)[:3], self.segments[k][2].new_ones(1)])
                               for k in g.segs]).contiguous()
            g.centres = cen
            g.seg_starts = torch.tensor([self.segments[k][0] - g.start for k in g.segs] + [g.n], dtype=torch.int64,
                                        device=self.dev)
        # the one device->host read: on a side stream into pinned memory, so that kernels the caller
        # queues on its own stream meanwhile (the similarity of forward_pass / bench) are not waited for
        main = torch.cuda.current_stream(self.dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        binned = torch.cuda.Event()
        binned.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(binned)
            self._sizes_host[:ng].copy_(self._class_scratch[:ng], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._side)
        self._pending_sizes = done
        self._binned = True
        self._classes_ready = False

    def _resolve_classes(self):
        """Wait for the class sizes read back by finalize_frontend() (normally long complete: the
        caller has queued the similarity kernels in the meantime) and turn them into launch offsets."""
        if not self._binned:
            self.finalize_frontend()
        if self._classes_ready:
            return self.max_count
        self._pending_sizes.synchronize()
        self._pending_sizes = None
        total = np.zeros((self.n_classes,), np.int64)
        for gi, g in enumerate(self.groups):
            sizes = self._sizes_host[gi, :self.n_classes].numpy().astype(np.int64)
            g.class_sizes = sizes
            off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
            g.class_offsets = (ctypes.c_int64 * (self.n_classes + 1))(*[int(v) for v in off])
            total += sizes
        self.class_sizes = total
        nz = np.nonzero(total[1:])[0]
        self.max_count = int(min(self.M, 128 * (int(nz[-1]) + 1))) if len(nz) else 1
        self._classes_ready = True
        return self.max_count

    def unit_work(self, unit):
        """Traversed voxels per `unit` consecutive rays of this rank (float64 device tensor): the weights of the
        work-balanced block plan of sharding.balanced_boundaries.  n_rays must be a multiple of unit."""
        c = self.count[:self.n_rays].to(torch.float64)
        return torch.where(c > 1, c, torch.zeros_like(c)).reshape(-1, int(unit)).sum(dim=1)

    # ------------------------------------------------------------------ BP
    def _group_rows(self, g):
        """(lin, s_hat) row tensors of a group: its own rows when resident, else the freshly scored window."""
        if g.resident:
            return self.lin[g.start:g.start + g.n], self.s_hat[g.start:g.start + g.n]
        k = g.segs[0]
        at = self.resident_capacity
        self._score(k, at, *self._scored[k])
        return self.lin[at:at + g.n], self.s_hat[at:at + g.n]

    def bp_iteration(self):
        if not self._classes_ready:
            self._resolve_classes()
        st = current_stream_ptr()
        if self._peer is None:
            self._fill(self.acc_new, sharding.seed_value(self.rank, self.prior))
        if self.sweep_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        first = self.iterations_done == 0
        for g in self.groups:
            if g.n == 0:
                continue
            sl = slice(g.start, g.start + g.n)
            if (g.resident and first and self._acc_uniform and self._unmapped and not self.parity
                    and sorted(self._unmapped) == list(g.segs)):
                # mapping fused into the first sweep: the rows are built where they are first used
                _lib.call("rn_engine_first_sweep_mapped", self.params, _ptr(self.axes), _ptr(self.starts[sl]),
                          _ptr(self.ends[sl]), _ptr(self.hdr[sl]), _ptr(self.codes[sl]), _ptr(self.count[sl]),
                          _ptr(self._planes_all[sl]), _ptr(self.lin[sl]), _ptr(self.s_hat[sl]), _ptr(self.msgs[sl]),
                          _ptr(self.acc_prev), _ptr(self.acc_new), _ptr(self.order[sl]), g.class_offsets, g.n, st)
                self.launches += int(np.count_nonzero(g.class_sizes))
                self._unmapped = []
                continue
            if g.resident:
                self._ensure_mapped()
            lin, s_hat = self._group_rows(g)
            if self.parity:
                _lib.call("rn_engine_bp_iteration_f64", self.params, _ptr(lin), _ptr(self.count[sl]), _ptr(s_hat),
                          _ptr(self.msgs[sl]), _ptr(self.acc_prev), _ptr(self.acc_new), 1 if first else 0, g.n, st)
                self.launches += 1
            else:
                _lib.call("rn_engine_bp_iteration", self.params, _ptr(lin), _ptr(self.count[sl]), _ptr(s_hat),
                          _ptr(self.msgs[sl]), _ptr(self.acc_prev), _ptr(self.acc_new), _ptr(self.order[sl]),
                          g.class_offsets, (2 if self._acc_uniform else 1) if first else 0, int(self.max_count), g.n, st)
                self.launches += int(np.count_nonzero(g.class_sizes[1:]))
        self._acc_uniform = False
        if self.sweep_events is not None:
            ev[1].record()
            self.sweep_events.append(ev)
        if self._peer is not None:
            # acc_prev (= the peer-mapped result buffer) <- prior + sum of partials; the other partial is cleared for
            # the next sweep and becomes acc_new
            if self.exchange_events is not None:
                xe = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                xe[0].record()
            self._peer.allreduce(self.prior)
            self.acc_new = self._peer.partial
            self.launches += 1
            if self.exchange_events is not None:
                xe[1].record()
                self.exchange_events.append(xe)
        else:
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
        """acc_prev as the reference's row-major [Gx, Gy, Gz] grid (float32; float64 in parity mode)."""
        out = torch.empty(self.grid_shape, dtype=self.acc_prev.dtype, device=self.dev)
        if self.parity:
            _lib.call("rn_bricks_to_grid_f64", self.params, _ptr(self.acc_prev), _ptr(out), None, current_stream_ptr())
        else:
            _lib.call("rn_bricks_to_grid", self.params, _ptr(self.acc_prev), _ptr(out), 0, current_stream_ptr())
        self.launches += 1
        return out

    def set_accumulator(self, grid):
        """Load a row-major [Gx, Gy, Gz] grid (CUDA tensor or numpy) as acc_prev."""
        dt = np.float64 if self.parity else np.float32
        if isinstance(grid, np.ndarray):
            grid = torch.from_numpy(np.ascontiguousarray(grid, dtype=dt)).to(self.dev)
        grid = grid.to(self.acc_prev.dtype).reshape(self.grid_shape).contiguous()
        _lib.call("rn_grid_to_bricks_f64" if self.parity else "rn_grid_to_bricks", self.params, _ptr(grid),
                  _ptr(self.acc_prev), self.prior, current_stream_ptr())
        self._acc_uniform = False
        self.launches += 1

    def set_messages(self, msgs):
        """Load messages [n_rays, <= M] and mark the state as past the first sweep."""
        if isinstance(msgs, np.ndarray):
            msgs = torch.from_numpy(np.ascontiguousarray(msgs, dtype=np.float32)).to(self.dev)
        self.msgs[:msgs.shape[0], :msgs.shape[1]].copy_(msgs)
        self.iterations_done = max(self.iterations_done, 1)

    def messages(self):
        """Messages [n_rays, M]; slots beyond a ray's count (and the rows of the rays BP skips) read 0 like the
        reference's zero-initialised array -- the kernels never touch them."""
        if self.iterations_done == 0:
            return torch.zeros((self.n_rays, self.M), dtype=torch.float32, device=self.dev)
        cnt = self.count[:self.n_rays]
        live = torch.arange(self.M, device=self.dev)[None, :] < torch.where(cnt > 1, cnt, torch.zeros_like(cnt))[:, None]
        return torch.where(live, self.msgs[:self.n_rays, :self.M], torch.zeros((), dtype=torch.float32, device=self.dev))

    # ------------------------------------------------------------------ outputs
    def depth(self, depth_out=None, S_new=None):
        """Depth per ray (flat, ray order of the trace_image calls); the resident segments in one launch."""
        if not self._classes_ready:
            self._resolve_classes()
        if depth_out is None:
            depth_out = torch.empty((self.n_rays,), dtype=torch.float32, device=self.dev)
        if self.iterations_done == 0:
            self.msgs[:self.n_rays].zero_()
        self._ensure_mapped()
        for g in self.groups:
            if g.n == 0:
                continue
            lin, s_hat = self._group_rows(g)
            sl = slice(g.start, g.start + g.n)
            _lib.call("rn_engine_depth_f64" if self.parity else "rn_engine_depth", self.params, _ptr(lin),
                      _ptr(self.count[sl]), _ptr(s_hat), _ptr(self.msgs[sl]), _ptr(self.acc_prev), _ptr(self.axes),
                      _ptr(g.centres), _ptr(g.seg_starts), len(g.segs), _ptr(depth_out[sl]),
                      _ptr(S_new[sl]) if S_new is not None else None, g.n, current_stream_ptr())
            self.launches += 1
        return depth_out

    def depth_distribution(self):
        """S_new [n_rays, M] (compute_depth_distribution, mrf_np.py:333-385) -- parity tests."""
        S_new = torch.empty((self.n_rays, self.R), dtype=torch.float32, device=self.dev)
        self.depth(S_new=S_new)
        return S_new[:, :self.M]

    def occupancy(self):
        """sigmoid(acc_prev) as a row-major float32 [Gx, Gy, Gz] grid (mrf_np.py:206-240)."""
        out = torch.empty(self.grid_shape, dtype=torch.float32, device=self.dev)
        if self.parity:
            _lib.call("rn_bricks_to_grid_f64", self.params, _ptr(self.acc_prev), None, _ptr(out), current_stream_ptr())
        else:
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
