"""tests/golden/scene_golden.npz: outputs of the reference's own scene helpers, executed in place
(raynet/utils/training_utils.py:9-68 get_adjacent_frames_idxs, common/parse_input_data.py:13-41
parse_scene_info, common/scene.py:232-257 camera files of tests/restrepo_mock_dataset/scene_1,
common/camera.py).  Run here:  python tests/golden/make_scene_golden.py"""
import os
import re

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def func_source(path, name, indent=""):
    src = open(os.path.join(REF, path)).read()
    start = src.index("%sdef %s(" % (indent, name))
    m = re.search(r"\n%s(def |class |@)" % indent, src[start + 1:])
    return src[start:start + 1 + m.start()] if m else src[start:]


def main():
    import xml.etree.ElementTree as ET
    ns = {"np": np, "ET": ET, "os": os}
    exec(func_source("raynet/utils/training_utils.py", "get_adjacent_frames_idxs"), ns)
    exec(func_source("raynet/common/parse_input_data.py", "parse_scene_info"), ns)
    cam_src = re.sub(r"^(from|import) .*$", "", open(os.path.join(REF, "raynet/common/camera.py")).read(), flags=re.M)
    exec(cam_src, ns)
    body = "class _S(object):\n" + func_source("raynet/common/scene.py", "_read_camera_poses", "    ")
    exec(body, ns)
    cases, outs = [], []
    for n_frames in (5, 12, 50):
        for n_adj in (2, 3, 4, 5):
            for skip in (0, 1):
                for ref in range(n_frames):
                    try:
                        r = np.asarray(ns["get_adjacent_frames_idxs"](ref, n_frames, n_adj, skip), np.int64)
                    except Exception:
                        continue
                    if len(r) != n_adj or (r < 0).any() or (r >= n_frames).any():
                        continue      # the reference wraps around in uint32 there: not a behaviour to pin
                    cases.append((ref, n_frames, n_adj, skip))
                    outs.append(np.pad(r, (0, 5 - n_adj), constant_values=-999))
    scene = os.path.join(REF, "tests/restrepo_mock_dataset/scene_1")
    bbox = ns["parse_scene_info"](os.path.join(scene, "scene_info.xml"))
    s = ns["_S"]()
    s._cam_paths = [os.path.join(scene, "cams_krt", f) for f in sorted(os.listdir(os.path.join(scene, "cams_krt")))]
    Ks, Rs, ts, Ps, Cs = [], [], [], [], []
    for i in range(len(s._cam_paths)):
        cp = s._read_camera_poses(i)
        cam = ns["Camera"](cp["K"], cp["R"], cp["t"])
        Ks.append(cp["K"]); Rs.append(cp["R"]); ts.append(cp["t"]); Ps.append(cam.P); Cs.append(cam.center)
    np.savez_compressed(os.path.join(HERE, "scene_golden.npz"), cases=np.array(cases), neighbors=np.array(outs), bbox=bbox,
                        K=np.stack(Ks), R=np.stack(Rs), t=np.stack(ts), P=np.stack(Ps), center=np.stack(Cs))
    print(len(cases), "neighbour cases;", bbox, len(Ks), "cameras")


if __name__ == "__main__":
    main()
