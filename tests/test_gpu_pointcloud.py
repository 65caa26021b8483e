"""depth-map fusion (rn_fuse_depth_maps / raynet_b200.pointcloud) against the fixture
produced by executing the reference's raynet/pointcloud.py (tests/golden/make_pointcloud_golden.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN_PC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pointcloud_golden.npz")


class _Cam(object):
    def __init__(self, P, P_pinv, centre):
        self.P, self.P_pinv, self.center = P, P_pinv, centre.reshape(4, 1)


class _Img(object):
    def __init__(self, cam):
        self.camera = cam


class _Scene(object):
    def __init__(self, g):
        self.images = [_Img(_Cam(g["P"][i], g["P_pinv"][i], g["centre"][i])) for i in range(g["P"].shape[0])]
        self.gt = g["gt"]

    def get_image(self, i):
        return self.images[i]

    def get_depth_map(self, i):
        return self.gt[i]


@pytest.mark.parametrize("borders,thr,nn", [(4, 0.05, 2), (0, 0.02, 3)])
def test_fusion_vs_reference_execution(borders, thr, nn, tmp_path):
    import torch
    assert torch.cuda.is_available()
    from raynet_b200.pointcloud import get_pointcloud
    g = np.load(GOLDEN_PC)
    key = "b%d_t%g_n%d" % (borders, thr, nn)
    scene = _Scene(g)
    frames = list(range(g["depth"].shape[0]))
    files = []
    for k in frames:                               # the reference passes .npy file names (pointcloud.py:131)
        p = str(tmp_path / ("d%d.npy" % k))
        np.save(p, g["depth"][k])
        files.append(p)
    plain = get_pointcloud(scene, frames, files, False, borders=borders)
    assert plain.points.shape == g["plain_" + key].shape
    assert np.abs(plain.points - g["plain_" + key]).max() <= 1e-6
    cons = get_pointcloud(scene, frames, list(g["depth"]), True, borders=borders, consistency_threshold=thr, n_neighbors=nn)
    assert np.array_equal(cons._neighbors(), g["neigh_" + key])
    assert cons.points.shape == g["cons_" + key].shape          # the same points survive the consistency check ...
    assert np.abs(cons.points - g["cons_" + key]).max() <= 1e-6  # ... in the same order
    assert np.isinf(cons.tau).sum() > 0 and (cons.tau[np.isfinite(cons.tau)] >= 0).all()
