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
import numpy as np
import torch

from .cuda_implementations.mvcnn_with_ray_marching_and_voxels_mapping import \
    batch_mvcnn_voxel_traversal_with_ray_marching_with_depth_estimation
from .cuda_implementations.raynet_fp import perform_raynet_fp
from .cuda_implementations.sample_points import compute_depth_from_distribution
from .cuda_implementations.similarities import perform_multi_view_cnn_forward_pass_with_depth_estimation
from .cuda_implementations.utils import device, to_gpu
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
        """forward_pass.py:168-179."""
        H, W = scene.image_shape
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
    forward_pass.py:622-641); per-ray state never leaves the device; when a torch.distributed
    process group is initialised, `images_range` selects THIS rank's reference images and the
    occupancy accumulator is all-reduced after every sweep (raynet_b200/sharding.py).

    Feature hook: a `model` that has `predict_features(scene, view_indices)` is asked for the
    (n, H+p+1, W+p+1, F) float32 maps of those views directly (e.g. a per-view feature cache, or
    pinned host memory); otherwise the reference's `model.predict(stack of zero-padded images)`
    is used."""

    def __init__(self, model, generation_params, sampling_scheme, image_shape, rays_batch, filter_out_rays=False,
                 bp_iterations=3):
        super(RayNetForwardPass, self).__init__(model, generation_params, sampling_scheme, image_shape,
                                                rays_batch, filter_out_rays)
        self.rays_batch = rays_batch
        self.ref_idx = -1
        self.bp_iterations = bp_iterations      # hard-coded to 3 in the reference (forward_pass.py:590)
        self.engine = None
        self._de = None
        self._feat_dev = None
        self._copy_stream = None
        self._staging = {}
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

    def _ray_ids(self, ray_idxs, n_pixels, dev, k=0):
        """Device int32 ray ids.  The usual case -- every pixel of the image is a ray
        (forward_pass.py:168-179 without filtering) -- is generated on the device once and cached;
        filtered ray sets are uploaded through a pinned buffer."""
        if len(ray_idxs) == n_pixels and (n_pixels == 0 or (int(ray_idxs[0]) == 0 and int(ray_idxs[-1]) == n_pixels - 1)):
            ids = self._staging.get("all_pixels")
            if ids is None or ids.shape[0] != n_pixels:
                ids = torch.arange(n_pixels, dtype=torch.int32, device=dev)
                self._staging["all_pixels"] = ids
            return ids
        host = self._pinned("ray_ids_%d" % k, (len(ray_idxs),), torch.int32)   # one buffer per image: copies are asynchronous
        host.numpy()[:] = ray_idxs
        self.h2d_bytes += host.numel() * 4
        return host.to(dev, non_blocking=True)

    def _make_engine(self, scene, F, n_rays_total):
        gp = self._generation_params
        vg = scene.voxel_grid(gp.grid_shape)
        M = int(gp.max_number_of_marched_voxels)
        eng = RayPotentialEngine(M, gp.depth_planes, gp.neighbors + 1, F, scene.image_shape[0],
                                 scene.image_shape[1], gp.padding, scene.bbox.ravel(), vg.shape[1:],
                                 gamma=gp.gamma_mrf if gp.gamma_mrf is not None else 0.05,
                                 max_rays=n_rays_total)
        eng.set_voxel_grid(vg)
        return eng

    def _view_features(self, scene, views):
        if hasattr(self._model, "predict_features"):
            f = self._model.predict_features(scene, views)
            if isinstance(f, torch.Tensor):
                return f
            return torch.from_numpy(np.ascontiguousarray(f, dtype=np.float32))
        chunks = []
        for i in range(0, len(views), 5):       # the reference predicts stacks of neighbors+1 = 5 images
            chunks.append(self._features([scene.get_image(v) for v in views[i:i + 5]]))
        return torch.from_numpy(np.concatenate(chunks, axis=0))

    def forward_pass(self, scene, images_range):
        assert isinstance(images_range, tuple)
        (start_img_idx, end_img_idx, skip) = images_range
        H, W = scene.image_shape
        dev = device()
        img_ids = list(range(start_img_idx, end_img_idx, skip))
        rays = [self.get_valid_rays_per_image(scene, i) for i in img_ids]
        total = int(sum(len(r) for r in rays))
        self.h2d_bytes = self.d2h_bytes = 0
        # ---- features: once per distinct view --------------------------------------------------
        # views are identified by object identity of the scene's cached Image objects
        # (common/scene.py:171-177 caches them per index)
        index_of = dict((id(scene.get_image(v)), v) for v in range(scene.n_images))
        orders = [[index_of[id(im)] for im in scene.get_image_with_neighbors(i)] for i in img_ids]
        views = sorted(set(v for o in orders for v in o))
        slot = dict((v, k) for k, v in enumerate(views))
        f_host = self._view_features(scene, views)
        main = torch.cuda.current_stream(dev)
        copied = {}       # view slot -> event after which its feature map is on the device
        if f_host.is_cuda:
            # the model produced the feature volume on the device (raynet_b200.models.SimpleCNN): nothing to upload
            self._feat_dev = f_host.contiguous()
            self.h2d_bytes += int(getattr(self._model, "last_h2d_bytes", 0))
        else:
            if self._feat_dev is None or self._feat_dev.shape != f_host.shape:
                self._feat_dev = torch.empty(f_host.shape, dtype=torch.float32, device=dev)
            # the feature maps travel on a copy stream, view by view in the order the reference images
            # need them, while the rays are traced and binned (neither needs them); the similarity of an
            # image waits only for the views it reads
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            self._copy_stream.wait_stream(main)      # the previous call's readers of _feat_dev are done
            first_use = []
            for o in orders:
                first_use += [slot[v] for v in o if slot[v] not in first_use]
            with torch.cuda.stream(self._copy_stream):
                for k in first_use:
                    self._feat_dev[k].copy_(f_host[k], non_blocking=True)
                    copied[k] = torch.cuda.Event()
                    copied[k].record()
            self.h2d_bytes += f_host.numel() * 4
        if self.engine is None or self.engine.capacity < total:
            self.engine = self._make_engine(scene, f_host.shape[-1], total)
        else:
            self.engine.reset()
        # ---- front end, first half: trace the rays of every reference image, bin them ----------
        # all camera matrices and view slots travel in ONE pinned buffer (pageable uploads block the
        # host for ~0.5 ms each and would leave the GPU idle between the per-image launches)
        n_img = len(img_ids)
        nV = len(orders[0])
        stride = 12 * nV + 16
        meta = self._pinned("cams", (n_img, stride), torch.float32)
        vids = self._pinned("view_ids", (n_img, nV), torch.int32)
        for k, ref_idx in enumerate(img_ids):
            images = scene.get_image_with_neighbors(ref_idx)
            assert len(images) == nV
            row = meta[k].numpy()
            row[:12 * nV] = np.array([im.camera.P for im in images], dtype=np.float32).ravel()
            row[12 * nV:12 * nV + 12] = np.asarray(images[0].camera.P_pinv, dtype=np.float32).ravel()
            row[12 * nV + 12:] = np.asarray(images[0].camera.center, dtype=np.float32).ravel()[:4]
            vids[k] = torch.tensor([slot[v] for v in orders[k]], dtype=torch.int32)
        meta_dev = meta.to(dev, non_blocking=True)
        vids_dev = vids.to(dev, non_blocking=True)
        self.h2d_bytes += meta.numel() * 4 + vids.numel() * 4
        per_image = []
        for k, ref_idx in enumerate(img_ids):
            ids = self._ray_ids(rays[k], H * W, dev, k)
            nP = 12 * nV
            self.engine.trace_image(ids, meta_dev[k, nP:nP + 12], meta_dev[k, nP + 12:nP + 16])
            per_image.append((meta_dev[k, :nP], vids_dev[k]))
        self.engine.finalize_frontend()
        self.d2h_bytes += 4
        # ---- front end, second half: similarity + plane->voxel mapping per reference image -----
        for k, (P_dev, view_ids) in enumerate(per_image):
            for v in orders[k]:
                if slot[v] in copied:
                    main.wait_event(copied.pop(slot[v]))
            self.engine.score_image(k, self._feat_dev, P_dev, view_ids=view_ids, n_feature_slots=len(views))
        # ---- BP sweeps + depth -----------------------------------------------------------------
        self.engine.run_bp(self.bp_iterations)
        depth_dev = self.engine.depth()
        depth_host = self._pinned("depth", (int(depth_dev.shape[0]),), torch.float32)
        depth_host.copy_(depth_dev, non_blocking=True)       # pinned destination: one DMA, no staging copy
        torch.cuda.current_stream(dev).synchronize()
        depth = depth_host.numpy().copy()
        self.d2h_bytes += depth.nbytes
        for k, ref_idx in enumerate(img_ids):
            start, n, _ = self.engine.segments[k]
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
