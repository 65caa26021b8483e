"""Per-scene forward-pass drivers -- the call surface of raynet/forward_pass.py.

    get_forward_pass_factory(name)(model, generation_params, sampling_scheme, image_shape,
                                   rays_batch[, filter_out_rays])
        .forward_pass(scene, (start, end, step))  ->  generator of float32 [H, W] depth maps

`model` is anything with `.predict(stack_of_zero_padded_images) -> (V, H+p+1, W+p+1, F)`
(the Keras MV-CNN in the reference, forward_pass.py:622; the CNN itself is out of scope).
`scene` needs image_shape, bbox, voxel_grid(grid_shape), get_image(i),
get_image_with_neighbors(i) (raynet/common/scene.py; raynet_b200.synth.SyntheticScene).

RayNetForwardPass implements the algorithm the reference intends (mrf_np.py:243-330; see
SURVEY.md 2.2 "latent defects" #3 for why the literal host loop of forward_pass.py:593-678
is not reproduced): features and the whole front end once per image, `bp_iterations`
synchronous sweeps over all rays of all images, then one depth pass per image.
"""
import time

import numpy as np
import torch

from .cuda_implementations.mvcnn_with_ray_marching_and_voxels_mapping import \
    batch_mvcnn_voxel_traversal_with_ray_marching_with_depth_estimation
from .cuda_implementations.raynet_fp import perform_raynet_fp
from .cuda_implementations.sample_points import compute_depth_from_distribution
from .cuda_implementations.similarities import perform_multi_view_cnn_forward_pass_with_depth_estimation
from .cuda_implementations.utils import device, to_gpu
from . import sharding
from .engine import RayPotentialEngine


class ForwardPass(object):
    """forward_pass.py:25-223."""

    def __init__(self, model, generation_params, sampling_scheme, image_shape, rays_batch=50000,
                 filter_out_rays=False):
        self._model = model
        self._generation_params = generation_params
        self._sampling_scheme = sampling_scheme
        self.rays_batch = rays_batch
        self._filter_out_rays = filter_out_rays
        self._fp = None

    @staticmethod
    def create_depth_map_from_distribution(scene, img_idx, S, truncate=800, sampling_scheme="sample_in_bbox"):
        """forward_pass.py:52-94."""
        H, W = scene.image_shape
        camera_center = scene.get_image(img_idx).camera.center
        D = compute_depth_from_distribution(
            np.arange(H * W, dtype=np.int32), scene.get_image(img_idx).camera.P_pinv, camera_center, H, W,
            scene.bbox.ravel(), S, np.arange(H * W, dtype=np.float32), sampling_scheme,
        ).reshape(W, H).T
        return np.minimum(D, truncate)

    def get_valid_rays_per_image(self, scene, i):
        """forward_pass.py:168-179 (the all-pixels index list is built once per image size and shared, read-only)."""
        H, W = scene.image_shape
        if not self._filter_out_rays:
            cached = getattr(self, "_all_pixels", None)
            if cached is None or cached.shape[0] != H * W:
                cached = self._all_pixels = np.arange(H * W, dtype=np.int32)
                cached.setflags(write=False)
            return cached
        idxs = np.arange(H * W, dtype=np.int32)
        if self._filter_out_rays:
            idxs = idxs.reshape(W, H).T
            G = scene.get_depth_map(i)
            return idxs[G != 0].ravel()
        return idxs

    def _to_list_with_zeropadded_images(self, images, inputs=None):
        """forward_pass.py:181-198."""
        if inputs is None:
            inputs = []
        H, W, C = images[0].image.shape
        p = self._generation_params.padding
        for im in images:
            zeropadded = np.zeros((H + 2 * p, W + 2 * p, C))
            zeropadded[p:p + H, p:p + W, :] = im.image
            inputs.append(zeropadded)
        return inputs

    def _features(self, images):
        return np.ascontiguousarray(
            self._model.predict(np.stack(self._to_list_with_zeropadded_images(images), axis=0)), dtype=np.float32)

    def forward_pass(self, scene, images_range):
        raise NotImplementedError()


class MultiViewCNNForwardPass(ForwardPass):
    """forward_pass.py:226-344: plane-sweep similarity + arg-max plane -> depth."""

    def __init__(self, model, generation_params, sampling_scheme, image_shape, rays_batch, filter_out_rays=False):
        super(MultiViewCNNForwardPass, self).__init__(model, generation_params, sampling_scheme, image_shape,
                                                      rays_batch, filter_out_rays)
        self.ref_idx = -1
        D = self._generation_params.depth_planes
        self.s_gpu = to_gpu(np.zeros((self.rays_batch, D), dtype=np.float32))
        self.points_gpu = to_gpu(np.zeros((self.rays_batch, D, 4), dtype=np.float32))

    def sim(self, scene, feature_size):
        if self._fp is None:
            self._fp = perform_multi_view_cnn_forward_pass_with_depth_estimation(
                self._generation_params.depth_planes, self._generation_params.neighbors + 1, feature_size,
                scene.image_shape[0], scene.image_shape[1], self._generation_params.padding,
                scene.bbox.ravel(), self._sampling_scheme)
        return self._fp

    def forward_pass(self, scene, images_range):
        assert isinstance(images_range, tuple)
        (start_img_idx, end_img_idx, skip) = images_range
        batch_size = self.rays_batch
        H, W = scene.image_shape
        self.ref_idx = start_img_idx
        while self.ref_idx < end_img_idx:
            ray_idxs = self.get_valid_rays_per_image(scene, self.ref_idx)
            images = scene.get_image_with_neighbors(self.ref_idx)
            features = self._features(images)
            features_gpu = to_gpu(features.ravel())
            ray_idxs_gpu = to_gpu(ray_idxs.astype(np.int32))
            P_gpu = to_gpu(np.array([im.camera.P for im in images], dtype=np.float32).ravel())
            P_inv_gpu = to_gpu(np.asarray(images[0].camera.P_pinv, dtype=np.float32).ravel())
            camera_center_gpu = to_gpu(np.asarray(images[0].camera.center, dtype=np.float32))
            F = features.shape[-1]
            depth_map = to_gpu(np.zeros((H * W), dtype=np.float32))
            for i in range(0, len(ray_idxs), batch_size):
                self.sim(scene, F)(ray_idxs_gpu[i:i + batch_size], features_gpu, P_gpu, P_inv_gpu,
                                   camera_center_gpu, self.s_gpu, self.points_gpu, depth_map[i:i + batch_size])
            self.ref_idx += skip
            yield depth_map.get().reshape(W, H).T


class MultiViewCNNVoxelSpaceForwardPass(ForwardPass):
    """forward_pass.py:347-485: similarity mapped to voxel space + arg-max voxel -> depth."""

    def __init__(self, model, generation_params, sampling_scheme, image_shape, rays_batch, filter_out_rays=False):
        super(MultiViewCNNVoxelSpaceForwardPass, self).__init__(model, generation_params, sampling_scheme,
                                                                image_shape, rays_batch, filter_out_rays)
        self.ref_idx = -1
        M = self._generation_params.max_number_of_marched_voxels
        self.s_gpu = to_gpu(np.zeros((rays_batch, M), dtype=np.float32))
        self.ray_voxel_count_gpu = to_gpu(np.zeros((rays_batch,), dtype=np.int32))
        self.ray_voxel_indices_gpu = to_gpu(np.zeros((rays_batch, M, 3), dtype=np.int32))
        self.voxel_grid_gpu = None

    def sim(self, scene, feature_size):
        if self._fp is None:
            grid_shape = np.array(scene.voxel_grid(self._generation_params.grid_shape).shape[1:])
            self._fp = batch_mvcnn_voxel_traversal_with_ray_marching_with_depth_estimation(
                self._generation_params.max_number_of_marched_voxels, self._generation_params.depth_planes,
                self._generation_params.neighbors + 1, feature_size, scene.image_shape[0], scene.image_shape[1],
                self._generation_params.padding, scene.bbox.ravel(), grid_shape, self._sampling_scheme)
        return self._fp

    def voxel_grid_to_gpu(self, scene):
        if self.voxel_grid_gpu is None:
            self.voxel_grid_gpu = to_gpu(
                scene.voxel_grid(self._generation_params.grid_shape).transpose(1, 2, 3, 0).ravel())
        return self.voxel_grid_gpu

    def forward_pass(self, scene, images_range):
        assert isinstance(images_range, tuple)
        (start_img_idx, end_img_idx, skip) = images_range
        batch_size = self.rays_batch
        H, W = scene.image_shape
        self.ref_idx = start_img_idx
        while self.ref_idx < end_img_idx:
            ray_idxs = self.get_valid_rays_per_image(scene, self.ref_idx)
            images = scene.get_image_with_neighbors(self.ref_idx)
            features = self._features(images)
            features_gpu = to_gpu(features.ravel())
            ray_idxs_gpu = to_gpu(ray_idxs.astype(np.int32))
            P_gpu = to_gpu(np.array([im.camera.P for im in images], dtype=np.float32).ravel())
            P_inv_gpu = to_gpu(np.asarray(images[0].camera.P_pinv, dtype=np.float32).ravel())
            camera_center_gpu = to_gpu(np.asarray(images[0].camera.center, dtype=np.float32))
            F = features.shape[-1]
            depth_map = to_gpu(np.zeros((H * W), dtype=np.float32))
            for i in range(0, len(ray_idxs), batch_size):
                self.s_gpu.fill(0)
                self.ray_voxel_indices_gpu.fill(0)
                self.ray_voxel_count_gpu.fill(0)
                self.sim(scene, F)(ray_idxs_gpu[i:i + batch_size], features_gpu, P_gpu, P_inv_gpu,
                                   camera_center_gpu, self.voxel_grid_to_gpu(scene), self.ray_voxel_indices_gpu,
                                   self.ray_voxel_count_gpu, self.s_gpu, depth_map[i:i + batch_size])
            self.ref_idx += skip
            yield depth_map.get().reshape(W, H).T


class RayNetForwardPass(ForwardPass):
    """forward_pass.py:488-748 on the resident engine.

    Differences from the literal host loop of the reference (all behaviour-preserving, see the
    module docstring): features are computed ONCE per distinct view and uploaded once (the
    reference re-runs the CNN and re-uploads a re-ordered copy per reference image per sweep,
    forward_pass.py:622-641); per-ray state never leaves the device.

    Memory.  The reference bounds its device memory with `rays_batch` (rays per launch) and spills
    the messages to disk (forward_pass.py:588,602-611).  Here `rays_batch` is accepted for signature
    compatibility and the bound is `memory_budget` (bytes; default 85 % of the free HBM): reference
    images whose per-ray state does not fit are STREAMED -- their messages stay resident, their
    voxel-space rows are recomputed on every sweep (engine.py) -- and a job that cannot fit even so
    raises MemoryError before anything is launched.

    Multi-GPU (a torch.distributed process group is initialised): `shard="rays"` (default) --
    `images_range` names the WHOLE job on every rank; the (image, column-major pixel) ray enumeration
    is cut into `world` contiguous blocks (sharding.image_segments), the occupancy accumulator is
    all-reduced after every sweep and every rank yields the complete depth maps.  `shard="images"` --
    `images_range` selects THIS rank's reference images (rank, n, world), only the accumulator is
    shared and a rank yields its own images' maps.  `shard="none"` -- ignore the process group (this process runs
    the whole job alone).

    Feature hook: a `model` that has `predict_features(scene, view_indices)` is asked for the
    (n, H+p+1, W+p+1, F) float32 maps of those views directly (raynet_b200.models.SimpleCNN leaves
    them on the device); otherwise the reference's `model.predict(stack of zero-padded images)` is
    used.  Under shard="rays" the views are dealt out to the ranks, each rank runs the model on its
    share and the maps are exchanged with one all-gather over NVLink."""

    def __init__(self, model, generation_params, sampling_scheme, image_shape, rays_batch, filter_out_rays=False,
                 bp_iterations=3, memory_budget=None, shard="rays", parity=False, collective="auto"):
        super(RayNetForwardPass, self).__init__(model, generation_params, sampling_scheme, image_shape,
                                                rays_batch, filter_out_rays)
        assert shard in ("rays", "images", "none")
        self.rays_batch = rays_batch
        self.ref_idx = -1
        self.bp_iterations = bp_iterations      # hard-coded to 3 in the reference (forward_pass.py:590)
        self.memory_budget = memory_budget
        self.shard = shard
        self.parity = parity
        self.collective = collective            # exchange step of the multi-GPU path: "auto", "peer" or "nccl" (engine.py)
        self.engine = None
        self._plans = {}                        # job -> work-balanced block boundaries (shard="rays")
        self._de = None
        self._feat_dev = None
        self._copy_stream = None
        self._staging = {}
        self._copy_pool = None
        self.profile = False                    # True: forward_pass() fills self.timings (ms per stage, synchronised)
        self.timings = {}
        self.h2d_bytes = 0                      # bytes copied host->device / device->host by the last
        self.d2h_bytes = 0                      # forward_pass() call (bench.py's e2e accounting)

    def raynet_fp(self, scene, feature_size):
        """The reference's per-batch closures (forward_pass.py:545-569), kept for callers that
        drive the batches themselves."""
        if self._fp is None:
            grid_shape = np.array(scene.voxel_grid(self._generation_params.grid_shape).shape[1:])
            self._fp, self._de = perform_raynet_fp(
                self._generation_params.max_number_of_marched_voxels, self._generation_params.depth_planes,
                self._generation_params.neighbors + 1, feature_size, scene.image_shape[0], scene.image_shape[1],
                self._generation_params.padding, scene.bbox.ravel(), grid_shape, self._sampling_scheme)
        return [self._fp, self._de]

    def _pinned(self, name, shape, dtype):
        """A cached pinned host staging buffer."""
        buf = self._staging.get(name)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(shape, dtype=dtype).pin_memory()
            self._staging[name] = buf
        return buf

    def _ray_ids(self, ray_idxs, first, last, n_pixels, dev, k=0):
        """Device int32 ray ids of ray_idxs[first:last].  The usual case -- every pixel of the image is
        a ray (forward_pass.py:168-179 without filtering) -- is generated on the device once and cached;
        filtered ray sets are uploaded through a pinned buffer."""
        if len(ray_idxs) == n_pixels and (n_pixels == 0 or (int(ray_idxs[0]) == 0 and int(ray_idxs[-1]) == n_pixels - 1)):
            ids = self._staging.get("all_pixels")
            if ids is None or ids.shape[0] != n_pixels:
                ids = torch.arange(n_pixels, dtype=torch.int32, device=dev)
                self._staging["all_pixels"] = ids
            return ids[first:last]
        host = self._pinned("ray_ids_%d" % k, (last - first,), torch.int32)   # one buffer per segment: copies are asynchronous
        host.numpy()[:] = ray_idxs[first:last]
        self.h2d_bytes += host.numel() * 4
        return host.to(dev, non_blocking=True)

    def _make_engine(self, scene, F, n_rays_total, max_segment):
        gp = self._generation_params
        vg = scene.voxel_grid(gp.grid_shape)
        M = int(gp.max_number_of_marched_voxels)
        eng = RayPotentialEngine(M, gp.depth_planes, gp.neighbors + 1, F, scene.image_shape[0],
                                 scene.image_shape[1], gp.padding, scene.bbox.ravel(), vg.shape[1:],
                                 gamma=gp.gamma_mrf if gp.gamma_mrf is not None else 0.05,
                                 max_rays=n_rays_total, parity=self.parity, memory_budget=self.memory_budget,
                                 max_segment_rays=max_segment, use_distributed=None if self.shard != "none" else False,
                                 collective=self.collective, fuse_first_sweep=True)
        eng.set_voxel_grid(vg)
        return eng

    def _predict_views(self, scene, views):
        if hasattr(self._model, "predict_features"):
            f = self._model.predict_features(scene, views)
            if isinstance(f, torch.Tensor):
                if f.is_cuda:
                    self.h2d_bytes += int(getattr(self._model, "last_h2d_bytes", 0))
                return f
            return torch.from_numpy(np.ascontiguousarray(f, dtype=np.float32))
        chunks = []
        for i in range(0, len(views), 5):       # the reference predicts stacks of neighbors+1 = 5 images
            chunks.append(self._features([scene.get_image(v) for v in views[i:i + 5]]))
        return torch.from_numpy(np.concatenate(chunks, axis=0))

    @staticmethod
    def _view_orders(scene, img_ids):
        """Scene indices of [reference, neighbours...] per reference image.  Scenes that know their
        neighbour indices say so (`view_order(i)`: synth.SyntheticScene, common.scene.Scene); otherwise
        the Image objects returned by get_image_with_neighbors are matched by identity against
        scene.get_image(j) for the images of this call first, then the rest of the scene."""
        if hasattr(scene, "view_order"):
            return [list(scene.view_order(i)) for i in img_ids]
        known = {}
        for j in img_ids:
            known[id(scene.get_image(j))] = j
        orders, keep = [], []
        for i in img_ids:
            ims = scene.get_image_with_neighbors(i)
            keep.append(ims)                     # keeps the objects alive: ids stay unique
            if any(id(im) not in known for im in ims):
                for j in range(scene.n_images):
                    im = scene.get_image(j)
                    keep.append(im)
                    known.setdefault(id(im), j)
            orders.append([known[id(im)] for im in ims])
        return orders

    def forward_pass(self, scene, images_range):
        assert isinstance(images_range, tuple)
        (start_img_idx, end_img_idx, skip) = images_range
        H, W = scene.image_shape
        dev = device()
        dist = torch.distributed
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized() and self.shard != "none") else 1
        rank = dist.get_rank() if world > 1 else 0
        by_rays = world > 1 and self.shard == "rays"
        img_ids = list(range(start_img_idx, end_img_idx, skip))
        stamps = []

        def stamp(name):        # self.profile = True: host wall clock per stage, the device drained at every stamp
            if self.profile:
                torch.cuda.synchronize(dev)
                stamps.append((name, time.perf_counter()))
        stamp("start")
        rays = [self.get_valid_rays_per_image(scene, i) for i in img_ids]
        self.h2d_bytes = self.d2h_bytes = 0
        # ---- this rank's segments: (position in img_ids, first ray, last ray) -------------------
        plan_key = plan_unit = None
        if by_rays:
            # first call on a job: blocks of equal ray counts; the traversed voxels per group of 8 image columns are
            # then summed over the ranks and later calls on the same job use blocks of equal WORK
            # (sharding.balanced_boundaries; rays near the image border cross few voxels)
            lens = [len(r) for r in rays]
            plan_unit = 8 * H if all(n % (8 * H) == 0 for n in lens) else 0
            plan_key = (id(scene), images_range, world, tuple(lens))
            bounds = self._plans.get(plan_key)
            if bounds is not None:
                all_segs = [sharding.segments_from_unit_boundaries(lens, plan_unit, bounds[r], bounds[r + 1])
                            for r in range(world)]
            else:
                unit = 64 * H if all(n % (64 * H) == 0 for n in lens) else 8 * H
                all_segs = [sharding.image_segments(lens, r, world, unit) for r in range(world)]
            segs = all_segs[rank]
            # the same on every rank, so that every rank decides alike whether the engine has to grow (its
            # construction is collective: peer-mapped accumulators)
            total_cap = max(sum(b - a for (_, a, b) in sg) for sg in all_segs)
        else:
            segs = [(k, 0, len(r)) for k, r in enumerate(rays)]
            total_cap = int(sum(b - a for (_, a, b) in segs))
        total = int(sum(b - a for (_, a, b) in segs))
        max_seg = max([b - a for (_, a, b) in segs] + [1])
        if by_rays:
            max_seg = max(max([b - a for (_, a, b) in sg] + [1]) for sg in all_segs)
        # ---- features: once per distinct view ---------------------------------------------------
        orders_all = self._view_orders(scene, img_ids)
        views = sorted(set(v for (k, _, _) in segs for v in orders_all[k]))
        slot = dict((v, k) for k, v in enumerate(views))
        main = torch.cuda.current_stream(dev)
        copied = {}       # view slot -> event after which its feature map is on the device
        first_use = []
        for (k, _, _) in segs:
            first_use += [slot[v] for v in orders_all[k] if slot[v] not in first_use]
        if by_rays and hasattr(self._model, "predict_features"):      # (collective: ranks without rays take part too)
            # every rank needs (nearly) every view: the views are dealt out round-robin (view j * world + r to rank r),
            # every rank runs the model on its share, and the maps are exchanged over NVLink: one all-gather per full
            # round of `world` views (equal chunks, no padding), broadcasts for the views of the last partial round
            all_views = sorted(set(v for o in orders_all for v in o))
            mine = all_views[rank::world]
            part = self._predict_views(scene, mine) if mine else None
            fshape = self._staging.get("feat_shape")
            if fshape is None:                   # ranks without a view learn the map shape from the others, once
                shapes = [None] * world
                dist.all_gather_object(shapes, tuple(part.shape[1:]) if part is not None else None)
                fshape = [s for s in shapes if s is not None][0]
                self._staging["feat_shape"] = fshape
            n_slots = len(all_views)
            if self._feat_dev is None or tuple(self._feat_dev.shape) != (n_slots,) + tuple(fshape):
                self._feat_dev = torch.empty((n_slots,) + tuple(fshape), dtype=torch.float32, device=dev)
            for j in range(len(mine)):
                self._feat_dev[j * world + rank].copy_(part[j].to(dev, non_blocking=True))
            full = n_slots // world
            for j in range(full):
                dist.all_gather_into_tensor(self._feat_dev[j * world:(j + 1) * world], self._feat_dev[j * world + rank])
            for v in range(full * world, n_slots):
                dist.broadcast(self._feat_dev[v], src=v - full * world)
            slot = dict((v, k) for k, v in enumerate(all_views))
        else:
            # a model that works on the device (raynet_b200.models.SimpleCNN) runs on the side stream: its image
            # upload and convolutions overlap with the tracing and binning of the rays, which need no features
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            self._copy_stream.wait_stream(main)
            with torch.cuda.stream(self._copy_stream):
                f_host = self._predict_views(scene, views)
                if f_host.is_cuda:
                    f_host = f_host.contiguous()
                    f_host.record_stream(main)
                    ready = torch.cuda.Event()
                    ready.record()
            n_slots = len(views)
            if f_host.is_cuda:
                self._feat_dev = f_host          # produced on the device: nothing to upload
                for k in range(n_slots):
                    copied[k] = ready
            else:
                if self._feat_dev is None or self._feat_dev.shape != f_host.shape:
                    self._feat_dev = torch.empty(f_host.shape, dtype=torch.float32, device=dev)
                # the feature maps travel on a copy stream, view by view in the order the reference images
                # need them, while the rays are traced and binned (neither needs them); the similarity of an
                # image waits only for the views it reads.  Pageable sources are staged through pinned memory
                # (a pageable non_blocking copy blocks the host and nothing would overlap).
                if not f_host.is_pinned():
                    stage = self._pinned("features", tuple(f_host.shape), torch.float32)
                    stage.copy_(f_host)
                    f_host = stage
                with torch.cuda.stream(self._copy_stream):      # (it already waits for the previous call's readers)
                    for k in first_use:
                        self._feat_dev[k].copy_(f_host[k], non_blocking=True)
                        copied[k] = torch.cuda.Event()
                        copied[k].record()
                self.h2d_bytes += f_host.numel() * 4
        stamp("features")
        F = int(self._feat_dev.shape[-1])
        if self.engine is None or self.engine.capacity < total_cap or self.engine.max_segment_rays < min(max_seg, total_cap):
            self.engine = None                     # release the old state before the larger one is allocated
            self.engine = self._make_engine(scene, F, total_cap, max_seg)
        else:
            self.engine.reset()
        # ---- front end, first half: trace the rays of every segment, bin them --------------------
        # all camera matrices and view slots travel in ONE pinned buffer (pageable uploads block the
        # host for ~0.5 ms each and would leave the GPU idle between the per-image launches)
        n_seg = len(segs)
        nV = len(orders_all[0])
        stride = 12 * nV + 16
        meta = self._pinned("cams", (max(n_seg, 1), stride), torch.float32)
        vids = self._pinned("view_ids", (max(n_seg, 1), nV), torch.int32)
        for j, (k, _, _) in enumerate(segs):
            images = scene.get_image_with_neighbors(img_ids[k])
            assert len(images) == nV
            row = meta[j].numpy()
            row[:12 * nV] = np.array([im.camera.P for im in images], dtype=np.float32).ravel()
            row[12 * nV:12 * nV + 12] = np.asarray(images[0].camera.P_pinv, dtype=np.float32).ravel()
            row[12 * nV + 12:] = np.asarray(images[0].camera.center, dtype=np.float32).ravel()[:4]
            vids[j] = torch.tensor([slot[v] for v in orders_all[k]], dtype=torch.int32)
        meta_dev = meta.to(dev, non_blocking=True)
        vids_dev = vids.to(dev, non_blocking=True)
        self.h2d_bytes += meta.numel() * 4 + vids.numel() * 4
        per_seg = []
        for j, (k, a, b) in enumerate(segs):
            ids = self._ray_ids(rays[k], a, b, H * W, dev, j)
            nP = 12 * nV
            self.engine.trace_image(ids, meta_dev[j, nP:nP + 12], meta_dev[j, nP + 12:nP + 16])
            per_seg.append((meta_dev[j, :nP], vids_dev[j]))
        self.engine.finalize_frontend()
        self.d2h_bytes += 4
        if by_rays and plan_unit and plan_key not in self._plans:
            n_units = int(sum(lens)) // plan_unit
            work = torch.zeros((n_units,), dtype=torch.float64, device=dev)
            if total:
                offs0 = np.concatenate([[0], np.cumsum(lens)])
                lo_u = int(offs0[segs[0][0]] + segs[0][1]) // plan_unit
                work[lo_u:lo_u + total // plan_unit] = self.engine.unit_work(plan_unit)
            dist.all_reduce(work, op=dist.ReduceOp.SUM)
            self._plans[plan_key] = sharding.balanced_boundaries(work.cpu().numpy(), world)
        # ---- front end, second half: similarity + plane->voxel mapping per segment ----------------
        for j, (P_dev, view_ids) in enumerate(per_seg):
            for v in orders_all[segs[j][0]]:
                if slot[v] in copied:
                    main.wait_event(copied.pop(slot[v]))
            self.engine.score_image(j, self._feat_dev, P_dev, view_ids=view_ids, n_feature_slots=n_slots)
        stamp("frontend")
        # ---- BP sweeps + depth -----------------------------------------------------------------
        self.engine.run_bp(self.bp_iterations)
        stamp("bp")
        if by_rays:
            # every rank yields complete maps: this rank's block goes into a zeroed job-sized buffer, one SUM
            # all-reduce (x + 0 is exact) completes it everywhere
            job_total = int(sum(len(r) for r in rays))
            offs = np.concatenate([[0], np.cumsum([len(r) for r in rays])])
            depth_dev = self._staging.get("depth_job")
            if depth_dev is None or depth_dev.shape[0] != job_total:
                depth_dev = torch.empty((job_total,), dtype=torch.float32, device=dev)
                self._staging["depth_job"] = depth_dev
            depth_dev.zero_()
            if segs:
                lo = int(offs[segs[0][0]] + segs[0][1])
                self.engine.depth(depth_out=depth_dev[lo:lo + total])
            dist.all_reduce(depth_dev, op=dist.ReduceOp.SUM)
            seg_of = [(int(offs[k]), len(rays[k])) for k in range(len(img_ids))]
        else:
            depth_dev = self.engine.depth()
            seg_of = [(self.engine.segments[k][0], self.engine.segments[k][1]) for k in range(len(img_ids))]
        depth_host = self._pinned("depth", (int(depth_dev.shape[0]),), torch.float32)
        depth_host.copy_(depth_dev, non_blocking=True)       # pinned destination: one DMA, no staging copy
        torch.cuda.current_stream(dev).synchronize()
        stamp("depth")
        # the pinned buffer is reused by the next call: hand out a private copy (four threads: a single memcpy of the
        # job's depth maps is the largest host-side item of a call on a multi-GPU box)
        src = depth_host.numpy()
        depth = np.empty_like(src)
        if src.shape[0] >= (1 << 20):
            if self._copy_pool is None:
                from concurrent.futures import ThreadPoolExecutor
                self._copy_pool = ThreadPoolExecutor(4)
            cuts = [src.shape[0] * i // 4 for i in range(5)]
            list(self._copy_pool.map(lambda ab: np.copyto(depth[ab[0]:ab[1]], src[ab[0]:ab[1]]), zip(cuts[:-1], cuts[1:])))
        else:
            np.copyto(depth, src)
        self.d2h_bytes += depth.nbytes
        stamp("copy_out")
        if self.profile:
            self.timings = dict((b[0], (b[1] - a[1]) * 1e3) for a, b in zip(stamps[:-1], stamps[1:]))
        for k, ref_idx in enumerate(img_ids):
            start, n = seg_of[k]
            if len(rays[k]) == H * W:
                d = depth[start:start + n]
            else:
                d = np.zeros((H * W,), dtype=np.float32)
                d[rays[k]] = depth[start:start + n]
            self.ref_idx = ref_idx
            yield d.reshape(W, H).T


def get_forward_pass_factory(name):
    """forward_pass.py:859-865 (the Hartmann et al. baseline is out of scope)."""
    return {
        "multi_view_cnn": MultiViewCNNForwardPass,
        "multi_view_cnn_voxel_space": MultiViewCNNVoxelSpaceForwardPass,
        "raynet": RayNetForwardPass,
    }[name]
