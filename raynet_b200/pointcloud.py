"""Depth maps -> point cloud, mirror of raynet/pointcloud.py:76-270.

`PointcloudFromDepthMaps` back-projects every kept pixel (inside the borders, ground truth != 0) of every
predicted depth map; `PointcloudFromDepthMapsWithConsistency` additionally drops a point when its
distance to any of the `n_neighbors` nearest cameras disagrees with that camera's predicted depth
map by `consistency_threshold` or more, or when it projects outside one of them.  The per-pixel work
runs on the GPU (`rn_fuse_depth_maps`, csrc/rn_fusion.cuh); `depthmaps` may be arrays or `.npy` file
names like in the reference.  `.points` is the reference's (3, N) float array.
"""
import numpy as np
import torch

from . import _lib
from .cuda_implementations.utils import current_stream_ptr, device


class Pointcloud(object):
    """pointcloud.py:14-31 (the KD-tree / PLY helpers are out of scope)."""

    def __init__(self, points):
        self._points = points

    @property
    def points(self):
        return self._points

    def save(self, file):
        np.save(file, self.points)


class PointcloudFromDepthMaps(Pointcloud):
    def __init__(self, scene, frame_idxs, depthmaps, borders=40):
        self._scene = scene
        self._frame_idxs = list(frame_idxs)
        self._depthmaps = list(depthmaps)
        self._borders = int(borders)
        self._points = None
        self.tau = None              # float32 [n, H, W] after .points was evaluated

    def _neighbors(self):
        return None

    def _threshold(self):
        return np.inf

    def _load(self, d):
        d = np.load(d) if isinstance(d, str) else np.asarray(d)
        d = np.array(d, dtype=np.float32)
        nan = np.isnan(d)
        if nan.any():                                  # pointcloud.py:133-135
            d[nan] = d[~nan].min()
        return d

    @property
    def points(self):
        if self._points is None:
            dev = device()
            depth = np.stack([self._load(d) for d in self._depthmaps])
            n, H, W = depth.shape
            cams = [self._scene.get_image(i).camera for i in self._frame_idxs]
            P = np.stack([np.asarray(c.P, np.float64) for c in cams])
            P_pinv = np.stack([np.asarray(c.P_pinv, np.float64) for c in cams])
            centre = np.stack([np.asarray(c.center, np.float64).ravel()[:4] for c in cams])
            gt = None       # ground-truth mask (pointcloud.py:118-121); a scene without ground truth keeps every pixel
            if hasattr(self._scene, "get_depth_map"):
                try:
                    gt = np.stack([np.asarray(self._scene.get_depth_map(i), np.float32) for i in self._frame_idxs])
                    assert gt.shape == depth.shape
                except NotImplementedError:
                    gt = None
            nb = self._neighbors()
            t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            d_depth, d_gt, d_P, d_Pi, d_C, d_nb = t(depth), t(gt), t(P), t(P_pinv), t(centre), t(nb)
            points = torch.empty((n, H, W, 3), dtype=torch.float32, device=dev)
            tau = torch.empty((n, H, W), dtype=torch.float32, device=dev)
            ptr = lambda x: None if x is None else x.data_ptr()
            _lib.call("rn_fuse_depth_maps", ptr(d_depth), ptr(d_gt), ptr(d_P), ptr(d_Pi), ptr(d_C), ptr(d_nb), n, H, W,
                      0 if nb is None else int(nb.shape[1]), self._borders, ptr(points), ptr(tau), current_stream_ptr())
            keep = tau < self._threshold()
            self.tau = tau.cpu().numpy()
            self._points = points[keep].T.contiguous().cpu().numpy()      # (3, N), images, then rows, then columns
        return self._points


class PointcloudFromDepthMapsWithConsistency(PointcloudFromDepthMaps):
    def __init__(self, scene, frame_idxs, depthmaps, borders=40, consistency_threshold=0.75, n_neighbors=5):
        super(PointcloudFromDepthMapsWithConsistency, self).__init__(scene, frame_idxs, depthmaps, borders)
        self._consistency_threshold = float(consistency_threshold)
        self._n_neighbors = int(n_neighbors)

    def _neighbors(self):
        """pointcloud.py:177-186, the reference's expression verbatim (positions in frame_idxs)."""
        a = np.hstack([np.asarray(self._scene.get_image(i).camera.center).reshape(4, 1) for i in self._frame_idxs])
        distances = 2 * (a * a).sum(axis=0) - 2 * (a.T.dot(a))
        return np.ascontiguousarray(distances.argsort()[:, 1:self._n_neighbors + 1].astype(np.int32))

    def _threshold(self):
        return self._consistency_threshold


def get_pointcloud(scene, frame_idxs, depthmaps, with_consistency, **kwargs):
    """pointcloud.py:248-270."""
    if with_consistency:
        return PointcloudFromDepthMapsWithConsistency(scene, frame_idxs, depthmaps, kwargs["borders"],
                                                      kwargs["consistency_threshold"], kwargs["n_neighbors"])
    return PointcloudFromDepthMaps(scene, frame_idxs, depthmaps, kwargs["borders"])
