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


@pytest.mark.parametrize("nq,nt,seed", [(5000, 20000, 0), (3000, 7, 1), (1000, 50000, 2)])
def test_nearest_neighbor_distances_vs_kdtree(nq, nt, seed):
    """rn_nn_grid_distances against the KD-tree query the reference uses (pointcloud.py:64-73; sklearn's
    KDTree here, Euclidean, k = 1): clustered targets, queries inside and well outside them."""
    from sklearn.neighbors import KDTree
    from raynet_b200.metrics import nearest_neighbor_distances
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-1, 1, size=(6, 3))
    target = (centres[rng.integers(0, 6, nt)] + rng.normal(0, 0.08, size=(nt, 3))).astype(np.float32)
    query = np.concatenate([rng.uniform(-1.5, 1.5, size=(nq // 2, 3)),
                            target[rng.integers(0, nt, nq - nq // 2)] + rng.normal(0, 0.01, size=(nq - nq // 2, 3))]).astype(np.float32)
    want = KDTree(target.astype(np.float64), 40, "minkowski").query(query.astype(np.float64), 1, True)[0].ravel()
    got = nearest_neighbor_distances(query.T, target.T)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-5 * max(1.0, want.max())


def test_accuracy_completeness_metrics():
    from sklearn.neighbors import KDTree
    from raynet_b200.metrics import Accuracy, Completeness
    from raynet_b200.pointcloud import Pointcloud
    rng = np.random.default_rng(5)
    gt = rng.uniform(-1, 1, size=(3, 4000)).astype(np.float32)
    pred = (gt[:, :3000] + rng.normal(0, 0.02, size=(3, 3000))).astype(np.float32)

    class Scene(object):
        def get_pointcloud(self):
            return Pointcloud(gt)

    acc, pts = Accuracy(truncate=0.03).compute(Scene(), [0], [None], Pointcloud(pred))
    want = np.minimum(KDTree(gt.T.astype(np.float64)).query(pred.T.astype(np.float64), 1)[0].ravel(), 0.03)
    assert pts is not None and np.abs(acc - want).max() <= 1e-6
    comp, _ = Completeness().compute(Scene(), [0], [None], Pointcloud(pred))
    want = KDTree(pred.T.astype(np.float64)).query(gt.T.astype(np.float64), 1)[0].ravel()
    assert np.abs(comp - want).max() <= 1e-5
