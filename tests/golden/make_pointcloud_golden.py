"""Generate tests/golden/pointcloud_golden.npz by EXECUTING the reference's depth-map fusion
(raynet/pointcloud.py:76-245 PointcloudFromDepthMaps / ...WithConsistency, common/camera.py:18-66,
common/image.py:242-258 rays(), utils/geometry.py:9-35 project) on a small synthetic rig.

Run here (needs /root/reference):  python tests/golden/make_pointcloud_golden.py
The reference modules are Python-2 era and import packages that are not installed (matplotlib,
its Cython fast_utils); their source is exec'ed in place after dropping the import lines and the
one py2 print statement -- every executed statement of the fusion is the reference's own.
"""
import os
import re
import sys
import tempfile
from itertools import product

import numpy as np

REF = "/root/reference/raynet"
HERE = os.path.dirname(os.path.abspath(__file__))


def ref_namespace():
    ns = {"np": np, "product": product, "sys": sys}
    def run(path, keep_from=None):
        src = open(os.path.join(REF, path)).read()
        src = re.sub(r"^(from|import) .*$", "", src, flags=re.M)                 # imports are provided by `ns`
        src = re.sub(r"^(\s*)print (.*)$", r"\1print(\2)", src, flags=re.M)     # py2 print statement
        exec(compile(src, path, "exec"), ns)
    # utils/geometry.py: only project() (the rest needs the Cython fast_utils)
    g = open(os.path.join(REF, "utils/geometry.py")).read()
    start = g.index("def project(")
    end = g.index("\ndef ", start + 1)
    exec(compile(g[start:end], "utils/geometry.py", "exec"), ns)
    run("common/camera.py")
    ns["KDTree"] = None
    ns["get_cmap"] = None
    run("pointcloud.py")
    # common/image.py: only Image.rays (the class needs imageio etc.)
    im = open(os.path.join(REF, "common/image.py")).read()
    start = im.index("    def rays(self):")
    end = im.index("\n    def ", start + 1) if "\n    def " in im[start + 1:] else len(im)
    body = "class _RaysMixin(object):\n" + im[start:end]
    exec(compile(body, "common/image.py", "exec"), ns)
    return ns


def main():
    ns = ref_namespace()
    Camera = ns["Camera"]
    rng = np.random.RandomState(0)
    n_img, H, W = 5, 36, 44
    # ring of pinhole cameras around the origin, like raynet_b200/synth.py
    cams = []
    for v in range(n_img):
        a = 2 * np.pi * v / n_img * 0.35          # neighbouring views overlap
        c = 3.0 * np.array([np.cos(a) * np.cos(0.5), np.sin(a) * np.cos(0.5), np.sin(0.5)])
        z = -c / np.linalg.norm(c)
        x = np.cross(z, [0, 0, 1.0]); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z])
        t = (-R.dot(c)).reshape(3, 1)
        f = 0.5 * W / np.tan(np.deg2rad(20))
        K = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1]])
        cams.append(Camera(K, R, t))

    class Img(ns["_RaysMixin"]):
        def __init__(self, cam):
            self._camera = cam
            self.camera = cam
            self.width, self.height = W, H

    images = [Img(c) for c in cams]
    # depth maps: distance to a sphere of radius 0.8 seen from each camera (smooth, consistent), plus noise
    depth, gt = [], []
    for k, im in enumerate(images):
        centre, rays = im.rays()
        d = rays - centre
        d = d[:3] / np.sqrt((d[:3] ** 2).sum(axis=0, keepdims=True))
        o = centre[:3]
        b = (o * d).sum(axis=0)
        disc = b ** 2 - ((o ** 2).sum() - 0.8 ** 2)
        hit = disc > 0
        tt = np.where(hit, -b - np.sqrt(np.maximum(disc, 0)), 3.5)
        D = tt.reshape(W, H).T.astype(np.float32)
        noise = rng.normal(0, 0.02, size=D.shape).astype(np.float32) * (rng.rand(*D.shape) < 0.3)
        depth.append((D + noise).astype(np.float32))
        G = hit.reshape(W, H).T.astype(np.float32)
        G[rng.rand(*G.shape) < 0.05] = 0
        gt.append(G)

    class Scene(object):
        def get_image(self, i):
            return images[i]
        def get_depth_map(self, i):
            return gt[i]

    tmp = tempfile.mkdtemp()
    files = []
    for k, D in enumerate(depth):
        p = os.path.join(tmp, "d%d.npy" % k)
        np.save(p, D)
        files.append(p)
    frame_idxs = list(range(n_img))
    out = {}
    for borders, thr, nn in ((4, 0.05, 2), (0, 0.02, 3)):
        plain = ns["PointcloudFromDepthMaps"](Scene(), frame_idxs, files, borders=borders)
        cons = ns["PointcloudFromDepthMapsWithConsistency"](Scene(), frame_idxs, files, borders=borders,
                                                            consistency_threshold=thr, n_neighbors=nn)
        key = "b%d_t%g_n%d" % (borders, thr, nn)
        out["plain_" + key] = plain.points
        out["cons_" + key] = cons.points
        out["neigh_" + key] = np.array([[j for j, _ in cons._neighbor_frames(i)] for i in frame_idxs], np.int32)
    np.savez_compressed(
        os.path.join(HERE, "pointcloud_golden.npz"),
        P=np.stack([c.P for c in cams]), P_pinv=np.stack([c.P_pinv for c in cams]),
        centre=np.stack([c.center.ravel() for c in cams]), depth=np.stack(depth), gt=np.stack(gt), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
