"""Synthetic camera rig + random feature volumes, shaped like the objects the
reference's forward passes consume (Scene / Image / Camera duck types).

Camera maths as raynet/common/camera.py:44-65 (P = K [R t], P+ = pinv(P), C = -R^-1 t, all
handed to the kernels as float32); voxel centres as raynet/utils/generic_utils.py:104-110.
"""
import numpy as np


class Camera(object):
    """Finite pinhole camera K, R, t (raynet/common/camera.py:4-65)."""

    def __init__(self, K, R, t):
        assert K.shape == (3, 3) and R.shape == (3, 3) and t.shape == (3, 1)
        self._K, self._R, self._t = K, R, t
        self._P = self._P_pinv = self._center = None

    @property
    def K(self):
        return self._K

    @property
    def R(self):
        return self._R

    @property
    def t(self):
        return self._t

    @property
    def center(self):
        if self._center is None:
            self._center = np.vstack([(-np.linalg.inv(self.R)).dot(self.t), [1]]).astype(np.float32)
        return self._center

    @property
    def P(self):
        if self._P is None:
            self._P = self._K.dot(np.hstack([self._R, self._t]))
        return self._P

    @property
    def P_pinv(self):
        if self._P_pinv is None:
            self._P_pinv = np.linalg.pinv(self.P)
        return self._P_pinv


class Image(object):
    def __init__(self, camera, image=None):
        self.camera = camera
        self.image = image


def get_voxel_grid(bbox, grid_shape):
    """(3, Gx, Gy, Gz) voxel centres, statement for statement utils/generic_utils.py:104-110."""
    bbox = np.asarray(bbox, dtype=np.float32).reshape(1, 6)
    xyz = [np.linspace(s, e, c, endpoint=False, dtype=np.float32)
           for s, e, c in zip(bbox[0, :3], bbox[0, 3:], grid_shape)]
    bin_size = np.array([xyzi[1] - xyzi[0] for xyzi in xyz]).reshape(3, 1, 1, 1)
    return (np.stack(np.meshgrid(*xyz, indexing="ij")) + bin_size / 2).astype(np.float32)


def ring_cameras(n_views, H, W, radius=3.0, elevation_deg=30.0, fov_deg=40.0):
    """V pinhole cameras on a ring looking at the origin, up = +z, f = 0.5 W / tan(fov/2)."""
    cams = []
    el = np.deg2rad(elevation_deg)
    f = 0.5 * W / np.tan(np.deg2rad(fov_deg) / 2.0)
    K = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1.0]])
    for v in range(n_views):
        a = 2.0 * np.pi * v / n_views
        C = radius * np.array([np.cos(a) * np.cos(el), np.sin(a) * np.cos(el), np.sin(el)])
        z = -C / np.linalg.norm(C)                 # optical axis: towards the origin
        up = np.array([0.0, 0.0, 1.0])
        x = np.cross(z, up)
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z])                    # world -> camera
        t = (-R.dot(C)).reshape(3, 1)
        cams.append(Camera(K, R, t))
    return cams


def random_features(n_views, H, W, F, padding, seed=0):
    """f32 [V, H+p+1, W+p+1, F], N(0,1)/sqrt(F); row 0 and column 0 (the "outside" slot of
    feature_similarities.cu:56-60) zeroed."""
    import torch
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    feats = torch.randn((n_views, H + padding + 1, W + padding + 1, F), generator=g, dtype=torch.float32)
    feats /= float(np.sqrt(F))
    feats[:, 0, :, :] = 0
    feats[:, :, 0, :] = 0
    return feats.numpy()


class SyntheticScene(object):
    """Duck type of raynet.common.scene.Scene for the forward passes: image_shape, bbox,
    voxel_grid(), get_image(), get_image_with_neighbors().  Every view is a reference view
    in turn with the next `neighbors` views of the ring (every `neighbor_stride`-th one) as its
    neighbours."""

    def __init__(self, n_views, H, W, grid_shape, bbox=(-1, -1, -1, 1, 1, 1), neighbors=None, with_images=False,
                 seed=0, neighbor_stride=1, pinned=None):
        self.n_views = n_views
        self._H, self._W = H, W
        self.grid_shape = np.asarray(grid_shape, dtype=np.int32)
        self._bbox = np.asarray(bbox, dtype=np.float32).reshape(1, 6)
        self.neighbors = (n_views - 1) if neighbors is None else neighbors
        self.neighbor_stride = int(neighbor_stride)
        cams = ring_cameras(n_views, H, W)
        rng = np.random.RandomState(seed)
        pixels = [None] * n_views
        if with_images:
            # the pixel buffers live in ONE block of (when a CUDA device is present) page-locked host memory,
            # like an image cache a data loader fills: uploads from it are true asynchronous DMAs
            import torch
            block = torch.empty((n_views, H, W, 3), dtype=torch.float32)
            if pinned is None:
                pinned = torch.cuda.is_available()
            if pinned:
                block = block.pin_memory()
            self._pixel_block = block
            pixels = block.numpy()
            for v in range(n_views):
                pixels[v] = rng.rand(H, W, 3).astype(np.float32)
        self.images = [Image(c, pixels[v]) for v, c in enumerate(cams)]
        self._voxel_grid = None

    @property
    def image_shape(self):
        return (self._H, self._W)

    @property
    def bbox(self):
        return self._bbox

    @property
    def n_images(self):
        return self.n_views

    def voxel_grid(self, grid_shape=None):
        if self._voxel_grid is None:
            self._voxel_grid = get_voxel_grid(self._bbox, self.grid_shape if grid_shape is None else grid_shape)
        return self._voxel_grid.astype(np.float32)

    def get_image(self, i):
        return self.images[i]

    def view_order(self, i):
        return [(i + k * self.neighbor_stride) % self.n_views for k in range(self.neighbors + 1)]

    def get_image_with_neighbors(self, i):
        return [self.images[j] for j in self.view_order(i)]


def camera_arrays(images):
    """(P f32[V,3,4], P_inv f32[4,3], centre f32[4]) of a [reference, neighbours...] list, cast
    exactly where the reference hands them to the device (forward_pass.py:631-641)."""
    P = np.stack([im.camera.P for im in images]).astype(np.float32)
    P_inv = np.asarray(images[0].camera.P_pinv, dtype=np.float32)
    centre = np.asarray(images[0].camera.center, dtype=np.float32).ravel()
    return P, P_inv, centre
