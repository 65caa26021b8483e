"""Shared synthetic cases for the parity tests.  Pure numpy; no GPU, no oracle."""
import numpy as np

from raynet_b200.synth import SyntheticScene, camera_arrays, get_voxel_grid, random_features

PADDING = 11
F = 32


class Case(object):
    """One reference image of the synthetic ring rig with everything the kernels consume."""

    def __init__(self, G, V, D, H, W, M, n_rays=None, ref_idx=0, seed=0, bbox=(-1, -1, -1, 1, 1, 1), grid=None,
                 feature_scale=3.0):
        self.G, self.V, self.D, self.H, self.W, self.M = G, V, D, H, W, M
        self.grid = np.asarray(grid if grid is not None else (G, G, G), np.int32)
        self.bbox = np.asarray(bbox, np.float32)
        self.scene = SyntheticScene(V, H, W, self.grid, bbox=bbox)
        # scaled up so that the plane softmax is peaked (scores ~ N(0, feature_scale^4 / F))
        self.features_all = random_features(V, H, W, F, PADDING, seed=seed) * np.float32(feature_scale)
        self.vgrid = np.ascontiguousarray(get_voxel_grid(self.bbox, self.grid).transpose(1, 2, 3, 0))
        self.set_reference(ref_idx, n_rays, seed)

    def set_reference(self, ref_idx, n_rays=None, seed=0):
        self.ref_idx = ref_idx
        order = self.scene.view_order(ref_idx)
        self.view_ids = np.asarray(order, np.int32)
        images = [self.scene.get_image(j) for j in order]
        self.P, self.P_inv, self.centre = camera_arrays(images)
        self.features = np.ascontiguousarray(self.features_all[order])             # reference's re-ordered copy
        if n_rays is None or n_rays >= self.H * self.W:
            self.ray_idxs = np.arange(self.H * self.W, dtype=np.int32)
        else:
            rng = np.random.default_rng(seed)
            self.ray_idxs = np.sort(rng.choice(self.H * self.W, size=n_rays, replace=False)).astype(np.int32)
        self.N = self.ray_idxs.shape[0]
        return self


def case_c1(**kw):
    """C1: 32^3 grid, 2 views, 16 planes, 1000 rays of a 64x64 image, M=96."""
    return Case(32, 2, 16, 64, 64, 96, n_rays=1000, **kw)


def case_small(**kw):
    """A 5-view / 32-plane case with rays long enough to span several 128-voxel chunks."""
    return Case(96, 5, 32, 48, 40, 288, n_rays=1200, **kw)


def case_long(**kw):
    """Rays of up to ~400 voxels (4 chunks of 128) on a 200^3 grid."""
    return Case(200, 3, 16, 32, 32, 600, n_rays=600, **kw)


def case_xlong(**kw):
    """C5's kernel parameters (15 views, 128 planes, M = 1536) on a small image: an elongated grid
    whose rays along x cross ~1100 voxels, i.e. 9-10 chunks of 128 (the longest length classes)."""
    return Case(None, 15, 128, 32, 32, 1536, n_rays=400, grid=(1000, 96, 96), **kw)


def case_nine(**kw):
    """Nine views / 64 planes (the headline's view and plane counts) on a small image and grid."""
    return Case(64, 9, 64, 40, 40, 192, n_rays=1000, **kw)


def case_dense(**kw):
    """SURVEY.md 7's probe: a 32^3 grid crossed by every pixel of 64x64 images from 4 views -- ~20 rays per
    voxel, accumulators far from 0, the setting in which BP amplifies a last-digit difference most."""
    return Case(32, 4, 16, 64, 64, 96, **kw)


def sigmoid(x):
    x = np.asarray(x, np.float64)
    return 1.0 / (1.0 + np.exp(-x))
