"""tests/golden/dtu_golden.npz: a tiny scene in the DTU MVS directory layout plus the outputs of the
reference's OWN DTUScene methods on it (raynet/common/scene.py:338-451 `_read_camera_poses`, `_get_depth_map`,
`get_depth_for_pixel`; common/parse_input_data.py:44-58 `parse_scene_info_dtu_dataset`), executed in place from
/root/reference with the imports of absent packages dropped.  The inputs are stored in the fixture so that the
test can rebuild the directory without the reference tree.   Run here:  python tests/golden/make_dtu_golden.py"""
import os
import re
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def func_source(path, name, indent=""):
    src = open(os.path.join(REF, path)).read()
    start = src.index("%sdef %s(" % (indent, name))
    m = re.search(r"\n%s(def |class |@)" % indent, src[start + 1:])
    return src[start:start + 1 + m.start()] if m else src[start:]


def write_scene(base, scan, K, Ps, pix, depths, bb):
    """The DTU layout for views 1..n of one scan (+ a view 50 and another illumination that must be ignored)."""
    from PIL import Image as PILImage
    from scipy.io import savemat
    rect = os.path.join(base, "Rectified", "scan%03d" % scan)
    cal = os.path.join(base, "SampleSet", "MVS_Data", "Calibration", "cal18")
    obs = os.path.join(base, "SampleSet", "MVS_Data", "ObsMask")
    dep = os.path.join(base, "Depth", "scan%03d" % scan)
    for d in (rect, cal, obs, dep):
        os.makedirs(d)
    for v in range(len(Ps)):
        PILImage.fromarray(pix[v]).save(os.path.join(rect, "rect_%03d_max.png" % (v + 1)))
        PILImage.fromarray(255 - pix[v]).save(os.path.join(rect, "rect_%03d_3_r5000.png" % (v + 1)))
        with open(os.path.join(cal, "pos_%03d.txt" % (v + 1)), "w") as f:
            f.write("\n".join(" ".join("%.9g" % x for x in row) for row in Ps[v]) + "\n")
        np.save(os.path.join(dep, "depth_%03d.npy" % (v + 1)), depths[v])
    PILImage.fromarray(pix[0]).save(os.path.join(rect, "rect_050_max.png"))
    with open(os.path.join(cal, "intrinsic.txt"), "w") as f:
        f.write("\n".join(" ".join("%.9g" % x for x in row) for row in K) + "\n0 0 0\n")
    savemat(os.path.join(obs, "ObsMask%d_10.mat" % scan), {"BB": bb, "ObsMask": np.ones((2, 2, 2), np.uint8)})


def make_inputs():
    from raynet_b200.synth import ring_cameras
    H, W, n = 12, 16, 4
    cams = ring_cameras(n, H, W, radius=600.0)
    K = cams[0].K.astype(np.float32)
    Ps = [c.P.astype(np.float32) for c in cams]
    rng = np.random.RandomState(3)
    pix = rng.randint(0, 256, size=(n, H, W, 3)).astype(np.uint8)
    depths = (500.0 + 200.0 * rng.rand(n, H, W)).astype(np.float32)
    depths[:, ::5, ::3] = 0
    bb = np.array([[-100.0, -120.0, -80.0], [110.0, 130.0, 90.0]])
    return H, W, K, Ps, pix, depths, bb


def main():
    from scipy.io import loadmat
    H, W, K, Ps, pix, depths, bb = make_inputs()
    base = tempfile.mkdtemp()
    write_scene(base, 7, K, Ps, pix, depths, bb)
    ns = {"np": np, "os": os, "loadmat": loadmat}
    exec(func_source("raynet/common/parse_input_data.py", "parse_scene_info_dtu_dataset"), ns)
    exec(func_source("raynet/utils/geometry.py", "project"), ns)
    exec(re.sub(r"^(from|import) .*$", "", open(os.path.join(REF, "raynet/common/camera.py")).read(), flags=re.M), ns)
    body = "class _S(object):\n" + "".join(func_source("raynet/common/scene.py", nm, "    ") for nm in ("_read_camera_poses", "_get_depth_map", "get_depth_for_pixel"))
    # the DTU methods are the LAST definitions of these names in scene.py
    src = open(os.path.join(REF, "raynet/common/scene.py")).read()
    dtu = src[src.index("class DTUScene(Scene):"):]
    def last(name):
        start = dtu.index("    def %s(" % name)
        m = re.search(r"\n    (def |@)", dtu[start + 1:])
        return dtu[start:start + 1 + m.start()] if m else dtu[start:]
    body = "class _S(object):\n" + last("_read_camera_poses") + "\n" + last("_get_depth_map") + "\n" + last("get_depth_for_pixel")
    body = body.replace("@lru_cache(maxsize=8)", "")
    ns["distance"] = lambda p1, p2: np.sqrt(np.sum((p1 - p2) ** 2))
    exec(body, ns)

    class Img(object):
        pass
    s = ns["_S"]()
    cal = os.path.join(base, "SampleSet", "MVS_Data", "Calibration", "cal18")
    s._cam_paths = [os.path.join(cal, f) for f in sorted(os.listdir(cal)) if "pos" in f]
    s._cam_intrinsic_path = os.path.join(cal, "intrinsic.txt")
    s._cache_depth_maps = [None] * len(Ps)
    images = []
    Ks, Rs, ts, Cs = [], [], [], []
    for i in range(len(Ps)):
        cp = s._read_camera_poses(i)
        cam = ns["Camera"](cp["K"], cp["R"], cp["t"])
        im = Img()
        im.camera = cam
        im.image = pix[i].astype(np.float32) / np.float32(255.)
        images.append(im)
        Ks.append(cp["K"]); Rs.append(cp["R"]); ts.append(cp["t"]); Cs.append(cam.center)
    s.get_image = lambda i: images[i]
    s.get_gt_depth_map = lambda i: depths[i]
    D = np.stack([s._get_depth_map(i) for i in range(len(Ps))])
    px = [(0, 3, 4), (1, 0, 0), (2, 7, 9), (3, 11, 15), (0, 5, 3)]
    dp = np.array([np.nan if s.get_depth_for_pixel(i, y, x) is None else s.get_depth_for_pixel(i, y, x) for (i, y, x) in px])
    bbox = ns["parse_scene_info_dtu_dataset"](os.path.join(base, "SampleSet", "MVS_Data", "ObsMask", "ObsMask7_10.mat"))
    np.savez_compressed(os.path.join(HERE, "dtu_golden.npz"), K_in=K, P_in=np.stack(Ps), pix=pix, depths=depths, bb=bb,
                        K=np.stack(Ks), R=np.stack(Rs), t=np.stack(ts), center=np.stack(Cs), depth_maps=D,
                        pixel_queries=np.array(px), pixel_depths=dp, bbox=bbox)
    print("views", len(Ps), "bbox", bbox, "depth maps", D.shape, "pixel depths", dp)


if __name__ == "__main__":
    main()
