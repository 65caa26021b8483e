"""Ray sharding across GPUs (one process per GPU) -- host-side logic only, no kernels.

The reference has no multi-GPU code.  The path shards because its BP is
synchronous (Jacobi): within a sweep every ray reads only the previous accumulator and adds
its messages into the new one (mrf_np.py:295-319), so rays can live on any rank as long as
the per-rank partial accumulators are summed once per sweep.

  * rays are enumerated in (reference image, column-major pixel) order and cut into `world`
    contiguous blocks, so each rank's rays stay spatially coherent;
  * every rank accumulates into a partial grid seeded with `seed_value(rank, prior)`: the
    prior on rank 0 and zero elsewhere, so that the SUM all-reduce directly yields
    prior + sum of all messages -- no epilogue pass over the grid;
  * `allreduce_accumulator` is the one collective of the path.

  * `PeerExchange` replaces that NCCL call on NVLink-connected GPUs by ONE kernel of this library over
    peer-mapped buffers (csrc/rn_peer.cuh): barrier, reduce + broadcast of one slice per rank, barrier, prior
    fused in.  torch.distributed._symmetric_memory only provides the peer mappings.

This module is what the world_size-2 gloo tests exercise on CPU.
"""
import ctypes

import torch


def ray_block(n_rays, rank, world):
    """Half-open [start, stop) of the contiguous block of `n_rays` owned by `rank`.
    Blocks differ by at most one ray; every ray is owned exactly once."""
    n_rays, rank, world = int(n_rays), int(rank), int(world)
    assert world >= 1 and 0 <= rank < world and n_rays >= 0
    base, extra = divmod(n_rays, world)
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    return start, stop


def aligned_ray_block(n_rays, rank, world, unit=1):
    """ray_block() with the interior boundaries rounded to multiples of `unit` rays (whole groups of
    image columns keep the tiled ray enumeration of the kernels usable inside partial images).
    Falls back to unit = 1 when n_rays is not a multiple of unit.  Every ray is owned exactly once."""
    n_rays, unit = int(n_rays), int(unit)
    if unit <= 1 or n_rays % unit != 0:
        return ray_block(n_rays, rank, world)
    a, b = ray_block(n_rays // unit, rank, world)
    return a * unit, b * unit


def image_segments(rays_per_image, rank, world, unit=1):
    """Split the concatenation of the images' ray lists into `world` contiguous blocks and
    return this rank's pieces as [(image_position, first, last_exclusive), ...] where first /
    last index into that image's ray list.  `rays_per_image`: list of ints.  unit: see
    aligned_ray_block (used only when every image is a multiple of it)."""
    total = int(sum(rays_per_image))
    if unit > 1 and any(int(n) % unit for n in rays_per_image):
        unit = 1
    start, stop = aligned_ray_block(total, rank, world, unit)
    out, off = [], 0
    for k, n in enumerate(rays_per_image):
        a, b = max(start, off), min(stop, off + n)
        if b > a:
            out.append((k, a - off, b - off))
        off += n
    return out


def balanced_boundaries(unit_weights, world):
    """Cut a sequence of work units (weights >= 0: e.g. traversed voxels per group of 8 image columns) into `world`
    contiguous blocks of (nearly) equal total weight.  Returns world + 1 non-decreasing unit indices, first 0, last
    len(unit_weights).  Deterministic: every rank computes the same plan from the same weights."""
    import numpy as np
    w = np.asarray(unit_weights, dtype=np.float64)
    n = int(w.shape[0])
    world = int(world)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    if total <= 0:
        return [ray_block(n, r, world)[0] for r in range(world)] + [n]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left"))
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, n)] - target):
            k -= 1                                  # the nearer of the two neighbouring cuts
        bounds.append(min(max(k, bounds[-1]), n))
    bounds.append(n)
    return bounds


def segments_from_unit_boundaries(rays_per_image, unit, lo_unit, hi_unit):
    """[(image_position, first, last_exclusive), ...] of the units [lo_unit, hi_unit) of the job's ray enumeration
    (every image a whole number of `unit` rays)."""
    out, off = [], 0
    start, stop = int(lo_unit) * int(unit), int(hi_unit) * int(unit)
    for k, n in enumerate(rays_per_image):
        a, b = max(start, off), min(stop, off + n)
        if b > a:
            out.append((k, a - off, b - off))
        off += n
    return out


def images_of_rank(n_images, rank, world):
    """Weak-scaling layout: whole reference images dealt out in contiguous runs."""
    start, stop = ray_block(n_images, rank, world)
    return list(range(start, stop))


def seed_value(rank, prior):
    """Initial value of a rank's partial accumulator (see module docstring)."""
    return float(prior) if int(rank) == 0 else 0.0


def allreduce_accumulator(acc, group=None):
    """Sum the per-rank partial accumulators in place (NCCL over NVLink on the GPUs; gloo in
    the CPU tests).  No-op outside a process group."""
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        if torch.distributed.get_world_size(group) > 1:
            torch.distributed.all_reduce(acc, op=torch.distributed.ReduceOp.SUM, group=group)
    return acc


class PeerExchange(object):
    """Peer-mapped partial / result accumulators of one rank + the fused exchange kernel.

        ex = PeerExchange(n_elements, device, group)      # collective: every rank of the group constructs it
        ex.partial   float32 [n]   the partial this rank's CURRENT sweep scatter-adds into (already zero)
        ex.result    float32 [n]   after ex.allreduce(prior): prior + sum over ranks of partial, on every rank
    The partials are double-buffered: allreduce() also clears the one the next sweep will use and flips
    `ex.partial` to it, so no fill is launched between sweeps; reset() zeroes the current one.
    Raises RuntimeError when the GPUs cannot map each other's memory (the caller then stays on NCCL)."""

    def __init__(self, n, device, group=None, n_ctas=None, multicast=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        from .cuda_implementations.utils import current_stream_ptr
        self._lib, self._stream = _lib, current_stream_ptr
        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n = int(n)
        assert self.n % 4 == 0
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        self.n_ctas = int(n_ctas or sms)      # (multicast kernel: 64 CTAs measured best, see below)
        try:
            self._partials = [symm_mem.empty(self.n, dtype=torch.float32, device=device) for _ in range(2)]
            self.result = symm_mem.empty(self.n, dtype=torch.float32, device=device)
            self.flags = symm_mem.empty(self.n_ctas * self.world, dtype=torch.int32, device=device)
            self.flags.zero_()
            for t in self._partials:
                t.zero_()
            handles = [symm_mem.rendezvous(t, group.group_name) for t in self._partials + [self.result, self.flags]]
        except Exception as e:      # no NVLink / P2P mapping between these devices
            raise RuntimeError("peer-mapped accumulators unavailable: %r" % (e,))
        self._handles = handles     # keep the mappings alive
        self._tables = [(ctypes.c_uint64 * self.world)(*[int(p) for p in h.buffer_ptrs]) for h in handles]
        # multicast (NVLS) addresses of the partials and the result: the switch then sums / broadcasts (rn_peer.cuh);
        # 0 where the fabric has no multicast support -- the peer-load kernel runs instead
        try:
            self._mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in handles[:3]]
        except Exception:
            self._mc = [0, 0, 0]
        # None: where it moves fewer bytes per link -- n (1 + 1 / world) against 2 n (world - 1) / world for peer loads,
        # measured for the 64 MiB grid of C3: 2 ranks 0.196 against 0.128 ms, 4 ranks 0.178 against 0.173 ms, 8 ranks
        # 0.174 against 0.221 ms (profiles/r02_exchange_microbench.json) -> beyond 4 ranks
        want = (self.world > 4) if multicast is None else bool(multicast)
        self.multicast = want and all(self._mc)
        if self.multicast and n_ctas is None:
            self.n_ctas = min(self.n_ctas, 64)      # 0.174 ms with 64 CTAs, 0.186 ms with 148 (8 GPUs, 64 MiB)
        self.cur = 0
        self.epoch = 0
        torch.cuda.synchronize(device)
        dist.barrier(group)         # every rank's flags and partials are zeroed before anyone signals

    @property
    def partial(self):
        return self._partials[self.cur]

    def allreduce(self, prior):
        nxt = self.cur ^ 1
        if self.multicast:
            self._lib.call("rn_peer_allreduce_mc_f32", self._mc[self.cur], self._mc[2], self._tables[3],
                           self._partials[nxt].data_ptr(), self.rank, self.world, self.n_ctas,
                           ctypes.c_uint32(self.epoch & 0xffffffff), float(prior), self.n, self._stream())
        else:
            self._lib.call("rn_peer_allreduce_f32", self._tables[self.cur], self._tables[2], self._tables[3],
                           self._partials[nxt].data_ptr(), self.rank, self.world, self.n_ctas,
                           ctypes.c_uint32(self.epoch & 0xffffffff), float(prior), self.n, self._stream())
        self.epoch += 2
        self.cur = nxt
        return self.result
