"""Accuracy / completeness of a predicted point cloud (second half): mirror of
raynet/metrics.py:156-236 on `nearest_neighbor_distances`, the GPU replacement of the reference's
sklearn KD-tree query (pointcloud.py:64-73)."""
import ctypes

import numpy as np
import torch

from . import _lib
from .cuda_implementations.utils import current_stream_ptr, device
from .pointcloud import Pointcloud, PointcloudFromDepthMaps


def nearest_neighbor_distances(query, target, points_per_cell=4.0, max_cells=1 << 24):
    """Distance of every column of `query` (3, Nq) to its nearest column of `target` (3, Nt): exact,
    Euclidean, float32 -- `KDTree(target.T).query(query.T, 1)[0]` of the reference."""
    dev = device()
    q = torch.as_tensor(np.ascontiguousarray(np.asarray(query, np.float32).T)).to(dev)
    t = torch.as_tensor(np.ascontiguousarray(np.asarray(target, np.float32).T)).to(dev)
    if t.shape[0] == 0:
        return np.full((q.shape[0],), np.inf, np.float32)
    lo = torch.minimum(t.min(dim=0).values, q.min(dim=0).values) if q.shape[0] else t.min(dim=0).values
    hi = torch.maximum(t.max(dim=0).values, q.max(dim=0).values) if q.shape[0] else t.max(dim=0).values
    ext = torch.clamp(hi - lo, min=1e-6).cpu().numpy().astype(np.float64)
    cell = float((np.prod(ext) * points_per_cell / max(int(t.shape[0]), 1)) ** (1.0 / 3.0))
    cell = max(cell, float(ext.max()) / 512.0)
    dims = np.maximum(np.ceil(ext / cell).astype(np.int64), 1)
    while int(np.prod(dims)) > max_cells:
        cell *= 1.26
        dims = np.maximum(np.ceil(ext / cell).astype(np.int64), 1)
    origin = lo.cpu().numpy().astype(np.float32)
    d_dims = torch.as_tensor(dims, device=dev)
    ijk = torch.floor((t - lo) / cell).long()
    ijk = torch.minimum(torch.clamp(ijk, min=0), d_dims - 1)
    cid = (ijk[:, 2] * int(dims[1]) + ijk[:, 1]) * int(dims[0]) + ijk[:, 0]
    order = torch.argsort(cid)
    t_sorted = t[order].contiguous()
    cell_start = torch.searchsorted(cid[order].contiguous(), torch.arange(int(np.prod(dims)) + 1, device=dev)).to(torch.int32)
    out = torch.empty((q.shape[0],), dtype=torch.float32, device=dev)
    o3 = (ctypes.c_float * 3)(*[float(v) for v in origin])
    d3 = (ctypes.c_int32 * 3)(*[int(v) for v in dims])
    _lib.call("rn_nn_grid_distances", q.data_ptr(), int(q.shape[0]), t_sorted.data_ptr(), cell_start.data_ptr(),
              ctypes.cast(o3, ctypes.c_void_p), ctypes.c_float(cell), ctypes.cast(d3, ctypes.c_void_p), 0, out.data_ptr(),
              current_stream_ptr())
    return out.cpu().numpy()


class _CloudMetric(object):
    def __init__(self, filter_factory=None, truncate=float("inf"), borders=40, use_pc_from_depthmap=False):
        self.filter_factory = filter_factory
        self.truncate = truncate
        self.borders = borders
        self.use_pc_from_depthmap = use_pc_from_depthmap

    def _ground_truth(self, scene, frame_idxs):
        if self.use_pc_from_depthmap:          # metrics.py:170-181
            return PointcloudFromDepthMaps(scene, frame_idxs, [scene.get_depthmap_file(i) for i in frame_idxs], self.borders)
        return scene.get_pointcloud()


class Accuracy(_CloudMetric):
    """metrics.py:156-195: distance of every predicted point to the ground-truth cloud, truncated."""

    def compute(self, scene, frame_idxs, depthmaps, predicted_pointcloud):
        gt = self._ground_truth(scene, frame_idxs)
        d = nearest_neighbor_distances(predicted_pointcloud.points, gt.points)
        return np.minimum(d, self.truncate), predicted_pointcloud.points


class Completeness(_CloudMetric):
    """metrics.py:198-236: distance of every ground-truth point to the predicted cloud, truncated."""

    def compute(self, scene, frame_idxs, depthmaps, predicted_pointcloud):
        gt = self._ground_truth(scene, frame_idxs)
        d = nearest_neighbor_distances(gt.points, predicted_pointcloud.points)
        return np.minimum(d, self.truncate), gt.points
