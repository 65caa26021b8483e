"""On-disk scenes for the forward passes: mirror of the parts of
raynet/common/scene.py, image.py and parse_input_data.py that `raynet_forward` needs to run on a
dataset in the Restrepo et al. layout

    <scene>/imgs/*            one image per view (sorted by file name)
    <scene>/cams_krt/*        per view: K (3 rows), blank, R (3 rows), blank, t (1 row)   (scene.py:232-257)
    <scene>/scene_info.xml    <bbox minx= miny= minz= maxx= maxy= maxz=>               (parse_input_data.py:13-41)
    <scene>/gt/gt_depth_%d.npy  optional ground-truth depth maps                          (scene.py:181-215)

and in the DTU MVS layout (scene.py:257-452)

    <base>/Rectified/scanNNN/rect_VVV_<illumination>.png      views 1..49 of scan NNN
    <base>/SampleSet/MVS_Data/Calibration/cal18/pos_VVV.txt   3x4 projection matrix per view
    <base>/SampleSet/MVS_Data/Calibration/cal18/intrinsic.txt K (first three rows)
    <base>/SampleSet/MVS_Data/ObsMask/ObsMask<N>_10.mat       "BB": bounding box, "ObsMask": observability mask
    <base>/Depth/scanNNN/*.npy                                ground-truth z-depth maps

Ground-truth meshes / STL point clouds, the octree ray caster and the sample generators stay out of scope.
"""
import os
import xml.etree.ElementTree as ET

import numpy as np

from ..synth import Camera, get_voxel_grid


def parse_scene_info(scene_info_filename):
    """1x6 float32 [minx miny minz maxx maxy maxz] (parse_input_data.py:13-41)."""
    attrib = dict((child.tag, child.attrib) for child in ET.parse(scene_info_filename).getroot())["bbox"]
    keys = ("minx", "miny", "minz", "maxx", "maxy", "maxz")
    return np.array([float(attrib[k]) for k in keys], dtype=np.float32).reshape(1, 6)


def parse_scene_info_dtu_dataset(scene_file):
    """1x6 float32 bounding box from the "BB" entry of a DTU ObsMask .mat file (parse_input_data.py:44-58)."""
    from scipy.io import loadmat
    return loadmat(scene_file, squeeze_me=True)["BB"].astype(np.float32).reshape(1, -1)


def project(P, point):
    """Affine transformation on homogeneous coordinates (utils/geometry.py:9-34): P (D1, D2) applied to the
    columns of point (D2, N), rows normalised by their last entry; one point comes back as a column vector."""
    points_hat = np.dot(P, point).T
    points_hat /= points_hat[:, -1:]
    if len(points_hat) == 1:
        points_hat = points_hat.T
    return points_hat


def get_adjacent_frames_idxs(ref_idx, n_frames, n_adjacent, skip):
    """Indices of the n_adjacent views around ref_idx, every (skip + 1)-th one, half before and half
    after where the sequence allows it and shifted inwards at its two ends -- the behaviour of
    utils/training_utils.py:9-68 (pinned by tests/golden/scene_golden.npz)."""
    if ref_idx > n_frames:
        raise ValueError("Ref index needs to be smaller than n_frames")
    step = skip + 1
    half = n_adjacent // 2
    first = max(0, ref_idx - half * step - (n_adjacent % 2))
    stop = min(n_frames, ref_idx + half * step + 1)
    idxs = [int(j) for j in range(first, ref_idx, step)] + [int(j) for j in range(ref_idx + 1, stop, step)]
    if len(idxs) != n_adjacent:
        if ref_idx == 0:
            idxs = list(range(step, (n_adjacent + 1) * step, step))
        elif ref_idx == n_frames - 1:
            idxs = list(range(ref_idx - n_adjacent * step, ref_idx, step))
        elif idxs and max(idxs) == n_frames - 1:      # ran into the end: extend towards the front
            while len(idxs) < n_adjacent:
                idxs.insert(0, min(idxs) - step)
        elif idxs and min(idxs) == 0:                 # ran into the front: extend towards the end
            while len(idxs) < n_adjacent:
                idxs.append(max(idxs) + step)
    return np.asarray(idxs, dtype=np.int64)


class Image(object):
    """image.py:11-70: pixel buffer (H, W, C) scaled to [0, 1] + the camera that produced it."""

    def __init__(self, camera, image_data, normalize=True):
        self._camera = camera
        image = np.asarray(image_data)
        if image.ndim == 2:
            image = image[:, :, np.newaxis]
        self._image = image.astype(np.float32) / np.float32(255.) if normalize else image

    @classmethod
    def from_file(cls, image_file, camera_poses):
        from PIL import Image as PILImage
        with PILImage.open(image_file) as im:
            data = np.array(im)
        return cls(Camera(K=camera_poses["K"], R=camera_poses["R"], t=camera_poses["t"]), data)

    image = property(lambda self: self._image)
    camera = property(lambda self: self._camera)
    width = property(lambda self: self._image.shape[1])
    height = property(lambda self: self._image.shape[0])
    channels = property(lambda self: self._image.shape[2])


class Scene(object):
    """scene.py:22-143: images + cameras + bounding box; neighbours by file order or camera distance."""

    def __init__(self, select_neighbors_based_on="filesystem"):
        self._voxel_grid = None
        self._camera_neighbors = None
        self._select_neighbors_based_on = select_neighbors_based_on

    @staticmethod
    def _load_sorted_files(basepath, directory, condition=None):
        path = os.path.join(basepath, directory)
        return [os.path.join(path, f) for f in sorted(filter(condition, os.listdir(path)))]

    def _get_neighbor_idxs(self, i, neighbors):
        if self._select_neighbors_based_on == "distance":
            if self._camera_neighbors is None:            # scene.py:58-76
                a = np.hstack([self.get_image(k).camera.center for k in range(self.n_images)])
                d = ((a.T[:, :, np.newaxis] - a[np.newaxis]) ** 2).sum(axis=1)
                self._camera_neighbors = d.argsort()[:, 1:neighbors + 1]
            return self._camera_neighbors[i]
        if self._select_neighbors_based_on == "filesystem":
            return get_adjacent_frames_idxs(i, self.n_images, neighbors, 0)
        raise NotImplementedError()

    @property
    def image_shape(self):
        im = self.get_image(0)
        return im.height, im.width

    def get_images(self):
        return [self.get_image(i) for i in range(self.n_images)]

    def view_order(self, i, neighbors=4):
        """Scene indices of [reference, neighbours...] -- what get_image_with_neighbors(i) returns, as indices
        (RayNetForwardPass keeps one feature map per distinct view)."""
        return [int(i)] + [int(n) for n in self._get_neighbor_idxs(i, neighbors)]

    def get_image_with_neighbors(self, i, neighbors=4):
        return [self.get_image(j) for j in self.view_order(i, neighbors)]

    def voxel_grid(self, grid_shape):
        if self._voxel_grid is None:
            if self.bbox is None:
                raise Exception("bbox needs to be different than None")
            self._voxel_grid = get_voxel_grid(self.bbox, grid_shape)
        return self._voxel_grid.astype(np.float32)


class RestrepoScene(Scene):
    """scene.py:144-268 without the ground-truth mesh machinery."""

    def __init__(self, basepath, select_neighbors_based_on="filesystem"):
        super(RestrepoScene, self).__init__(select_neighbors_based_on)
        self._basepath = basepath
        self._image_paths = self._load_sorted_files(basepath, "imgs")
        self._cam_paths = self._load_sorted_files(basepath, "cams_krt")
        self._bbox_path = os.path.join(basepath, "scene_info.xml")
        self._bbox = None
        self._cache = [None] * len(self._image_paths)

    @property
    def n_images(self):
        return len(self._image_paths)

    @property
    def bbox(self):
        if self._bbox is None:
            self._bbox = parse_scene_info(self._bbox_path)
        return self._bbox

    def get_image(self, i):
        if self._cache[i] is None:
            self._cache[i] = Image.from_file(self._image_paths[i], self._read_camera_poses(i))
        return self._cache[i]

    def _read_camera_poses(self, i):
        with open(self._cam_paths[i]) as f:
            rows = [line.split() for line in f if line.strip()]
        return {"K": np.array(rows[0:3]).astype(np.float32), "R": np.array(rows[3:-1]).astype(np.float32),
                "t": np.array(rows[-1]).astype(np.float32).reshape(-1, 1)}

    def get_depthmap_file(self, i):
        f = os.path.join(self._basepath, "gt", "gt_depth_%d.npy" % (i,))
        return f if os.path.isfile(f) else None

    def get_depth_map(self, i):
        f = self.get_depthmap_file(i)
        if f is None:
            raise NotImplementedError("no ground-truth depth map for view %d (the mesh ray caster is out of scope)" % i)
        return np.load(f)


class DTUScene(Scene):
    """scene.py:257-452: one scan of the DTU MVS dataset (views 1..49 under one illumination setting)."""

    def __init__(self, basepath, scene_idx, illumination="max", select_neighbors_based_on="filesystem"):
        super(DTUScene, self).__init__(select_neighbors_based_on)
        self._basepath = basepath
        self._image_paths = self._load_sorted_files(basepath, os.path.join("Rectified", "scan%03d" % (scene_idx,)),
                                                    lambda f: illumination in f)
        # ground-truth depth maps exist for the first 49 frames only (scene.py:276-284)
        self._image_paths = [ip for ip in self._image_paths
                             if int(ip.split("/")[-1].split(".")[0].split("_")[1]) <= 49]
        cal = os.path.join("SampleSet", "MVS_Data", "Calibration", "cal18")
        self._cam_paths = self._load_sorted_files(basepath, cal, lambda f: "pos" in f)
        self._cam_intrinsic_path = os.path.join(basepath, cal, "intrinsic.txt")
        self._bbox_path = os.path.join(basepath, "SampleSet", "MVS_Data", "ObsMask", "ObsMask%d_10.mat" % (scene_idx,))
        depth_dir = os.path.join("Depth", "scan%03d" % (scene_idx,))
        self._depth_map_paths = (self._load_sorted_files(basepath, depth_dir, lambda f: f.endswith("npy"))
                                 if os.path.isdir(os.path.join(basepath, depth_dir)) else [])
        self._bbox = None
        self._cache = [None] * len(self._image_paths)
        self._cache_depth_maps = [None] * len(self._image_paths)

    @property
    def n_images(self):
        return len(self._image_paths)

    @property
    def bbox(self):
        if self._bbox is None:
            self._bbox = parse_scene_info_dtu_dataset(self._bbox_path).astype(np.float32)
        return self._bbox

    @property
    def observation_mask(self):
        from scipy.io import loadmat
        return loadmat(self._bbox_path)["ObsMask"]

    def get_image(self, i):
        if self._cache[i] is None:
            self._cache[i] = Image.from_file(self._image_paths[i], self._read_camera_poses(i))
        return self._cache[i]

    def _read_camera_poses(self, i):
        """K from intrinsic.txt, [R | t] = K^-1 P with P the view's 3x4 projection matrix (scene.py:338-375)."""
        with open(self._cam_intrinsic_path) as f:
            rows = [x.strip().split(" ") for x in f.readlines()]
        K = np.array(rows[0:3]).astype(np.float32)
        with open(self._cam_paths[i]) as f:
            rows = [x.strip().split(" ") for x in f.readlines()]
        P = np.array(rows[0:4]).astype(np.float32)
        Rt = np.dot(np.linalg.inv(K), P)
        return {"K": K, "R": Rt[:, :3], "t": Rt[:, -1].reshape(-1, 1)}

    def get_gt_depth_map(self, i):
        return np.load(self._depth_map_paths[i])

    def get_depth_map(self, i):
        """Ground-truth z-depth -> distance from the camera centre per pixel, 0 where there is no ground truth
        (scene.py:382-416)."""
        if self._cache_depth_maps[i] is None:
            image = self.get_image(i)
            gt = self.get_gt_depth_map(i)
            H, W, _ = image.image.shape
            pixels = np.array([[u, v, 1.] for u in range(W) for v in range(H)], dtype=np.float32).T
            p_cc = np.dot(np.linalg.inv(image.camera.K), pixels)
            p_cc = p_cc * gt.T.reshape(1, -1)
            p_cc = np.vstack([p_cc, np.ones(p_cc.shape[1], dtype=np.float32)])
            P = np.vstack([np.hstack([image.camera.R, image.camera.t]), np.array([0., 0., 0., 1.])])
            target = project(np.linalg.inv(P), p_cc)
            D = np.sqrt(((target - image.camera.center.T) ** 2).sum(axis=-1)).reshape(W, H).T
            D *= (gt != 0)
            self._cache_depth_maps[i] = D.astype(np.float32)
        return self._cache_depth_maps[i]

    def get_depth_for_pixel(self, i, y, x):
        """Distance from the camera centre of the ground-truth surface point behind pixel (y, x), None without
        ground truth (scene.py:421-451)."""
        depth_value = self.get_gt_depth_map(i)[y, x]
        if depth_value == 0:
            return None
        im = self.get_image(i)
        p_cc = np.dot(np.linalg.inv(im.camera.K), np.array([[x, y, 1]], dtype=np.int32).T) * depth_value
        p_cc = np.vstack((p_cc, np.array([1])))
        P = np.vstack([np.hstack([im.camera.R, im.camera.t]), np.array([0., 0., 0., 1.])])
        target = project(np.linalg.inv(P), p_cc)
        return float(np.sqrt(np.sum((target[:-1] - im.camera.center[:-1]) ** 2)))

    def get_pointcloud(self):
        raise NotImplementedError("the ground-truth STL point clouds of DTU are outside the hot path (scene.py:453-455)")
