"""the MV-CNN feature extractor (rn_conv3x3_bn_relu / raynet_b200.models.SimpleCNN)
against the CPU restatement of the reference's Keras model (oracle/cnn_np.py, models.py:90-111)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.mark.parametrize("tensor_cores", [False, True])
@pytest.mark.parametrize("channels,shape", [(3, (2, 37, 45)), (1, (1, 30, 75)), (3, (3, 11, 11)), (3, (1, 64, 97)),
                                            (3, (2, 150, 300))])
def test_simple_cnn_vs_oracle(torch_cuda, channels, shape, tensor_cores, parity_log):
    """Both code paths of the 32 -> 32 layers: fp32 on the CUDA cores and 3 x TF32 on tcgen05 (rn_cnn_tc.cuh; the
    last shape spans several 128-pixel strips with a partial one and several 32-row chunks with a partial one)."""
    from oracle import cnn_np
    from raynet_b200.models import SimpleCNN
    rng = np.random.default_rng(3)
    model = SimpleCNN.random_init(channels=channels, seed=5, tensor_cores=tensor_cores)
    X = rng.uniform(0, 1, size=shape + (channels,)).astype(np.float32)
    got = model.predict(X)
    ref = cnn_np.simple_cnn_forward(X, model.get_weights())
    assert got.shape == ref.shape == (shape[0], shape[1] - 10, shape[2] - 10, 32)
    scale = np.abs(ref).max()
    err = np.abs(got - ref).max()
    print("cnn %s tensor_cores=%s: max |features - oracle| = %.2e (max |feature| %.2f)" % (shape, tensor_cores, err, scale))
    parity_log["cnn/%s/%s" % ("tcgen05_3xtf32" if tensor_cores else "cuda_cores_fp32", "x".join(map(str, shape)))] = {
        "max_abs_err_vs_float64_oracle": float(err), "max_abs_feature": float(scale)}
    assert err <= 1e-5 * max(scale, 1.0)
    assert model.launches == 5


def test_single_layer_entry_point_and_errors(torch_cuda):
    torch = torch_cuda
    from oracle import cnn_np
    from raynet_b200 import _lib
    rng = np.random.default_rng(0)
    x = rng.normal(size=(1, 20, 41, 32)).astype(np.float32)
    k = (rng.normal(size=(3, 3, 32, 32)) * 0.1).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, 32).astype(np.float32)
    shift = rng.normal(size=32).astype(np.float32)
    d = [torch.from_numpy(a).cuda() for a in (x, k, scale, shift)]
    out = torch.empty((1, 18, 39, 32), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.call("rn_conv3x3_bn_relu", d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), out.data_ptr(),
              1, 20, 41, 32, 0, st)
    ref = cnn_np.conv3x3_valid(x, k, np.zeros(32)) * scale + shift
    assert np.abs(out.cpu().numpy() - ref).max() <= 2e-5
    _lib.call("rn_conv3x3_bn_relu", d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), out.data_ptr(),
              1, 20, 41, 32, 1, st)
    assert np.abs(out.cpu().numpy() - np.maximum(ref, 0)).max() <= 2e-5
    with pytest.raises(NotImplementedError):
        _lib.call("rn_conv3x3_bn_relu", d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), out.data_ptr(),
                  1, 20, 41, 7, 0, st)
    with pytest.raises(AssertionError):
        _lib.call("rn_conv3x3_bn_relu", d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), out.data_ptr(),
                  1, 2, 41, 32, 0, st)


def test_raynet_forward_pass_from_images(torch_cuda, oracle):
    """images -> SimpleCNN -> resident ray-potential pipeline -> depth maps, through the reference's
    generator API; the CNN's two calling conventions (Keras-style predict on host stacks, and the
    predict_features hook that keeps the feature volume on the device) must give the same depth maps,
    and the device features must equal the oracle CNN's."""
    from oracle import cnn_np
    from raynet_b200.common.generation_parameters import GenerationParameters
    from raynet_b200.forward_pass import get_forward_pass_factory
    from raynet_b200.models import SimpleCNN
    from raynet_b200.synth import SyntheticScene
    V, H, W, G, D, M = 3, 24, 20, 24, 8, 72
    scene = SyntheticScene(V, H, W, (G, G, G), with_images=True)
    gp = GenerationParameters(depth_planes=D, neighbors=V - 1, grid_shape=np.array([G, G, G], np.int32),
                              max_number_of_marched_voxels=M, padding=11, gamma_mrf=0.05)
    cnn = SimpleCNN.random_init(channels=3, seed=1)

    class KerasLike(object):        # only the reference's model.predict(X) entry point
        def predict(self, X):
            return cnn.predict(X)

    maps = {}
    for name, model in (("hook", cnn), ("predict", KerasLike())):
        fp = get_forward_pass_factory("raynet")(model, gp, "sample_in_bbox", scene.image_shape, H * W)
        maps[name] = list(fp.forward_pass(scene, (0, V, 1)))
        assert len(maps[name]) == V and all(m.shape == (H, W) and np.isfinite(m).all() for m in maps[name])
    for a, b in zip(maps["hook"], maps["predict"]):
        assert np.array_equal(a, b)
    feats = cnn.predict_features(scene, [0, 1, 2]).cpu().numpy()
    X = np.zeros((V, H + 22, W + 22, 3), np.float32)
    for v in range(V):
        X[v, 11:11 + H, 11:11 + W] = scene.get_image(v).image
    ref = cnn_np.simple_cnn_forward(X, cnn.get_weights())
    assert feats.shape == (V, H + 12, W + 12, 32) and np.abs(feats - ref).max() <= 1e-5


def test_dataset_on_disk_to_point_cloud(torch_cuda, tmp_path):
    """SURVEY.md 8(f) rows 4 + 1 + path + 2 in one go: a scene in the Restrepo layout on disk ->
    RestrepoScene -> SimpleCNN features -> ray-potential inference -> depth maps -> fused point cloud."""
    import os
    from PIL import Image as PILImage
    from raynet_b200.common.generation_parameters import GenerationParameters
    from raynet_b200.common.scene import RestrepoScene
    from raynet_b200.forward_pass import get_forward_pass_factory
    from raynet_b200.models import SimpleCNN
    from raynet_b200.pointcloud import get_pointcloud
    from raynet_b200.synth import ring_cameras
    H, W, n, G, D, M = 24, 32, 6, 24, 8, 72
    os.makedirs(tmp_path / "imgs")
    os.makedirs(tmp_path / "cams_krt")
    rng = np.random.RandomState(1)
    for k, c in enumerate(ring_cameras(n, H, W)):
        PILImage.fromarray(rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)).save(str(tmp_path / "imgs" / ("f%03d.png" % k)))
        with open(str(tmp_path / "cams_krt" / ("f%03d_cam.txt" % k)), "w") as f:
            for Mx in (c.K, c.R):
                f.write("\n".join(" ".join("%.9g" % v for v in row) for row in Mx) + "\n\n")
            f.write(" ".join("%.9g" % v for v in c.t.ravel()) + "\n")
    with open(str(tmp_path / "scene_info.xml"), "w") as f:
        f.write('<bwm_info_for_boxm2><bbox maxx="1" maxy="1" maxz="1" minx="-1" miny="-1" minz="-1"></bbox></bwm_info_for_boxm2>')
    scene = RestrepoScene(str(tmp_path))
    gp = GenerationParameters(depth_planes=D, neighbors=4, grid_shape=np.array([G, G, G], np.int32),
                              max_number_of_marched_voxels=M, padding=11, gamma_mrf=0.05)
    fp = get_forward_pass_factory("raynet")(SimpleCNN.random_init(3, seed=2), gp, "sample_in_bbox", scene.image_shape, H * W)
    maps = list(fp.forward_pass(scene, (0, n, 1)))
    assert len(maps) == n and all(m.shape == (H, W) and np.isfinite(m).all() and (m > 0).all() for m in maps)
    pc = get_pointcloud(scene, list(range(n)), maps, True, borders=2, consistency_threshold=0.5, n_neighbors=2)
    pts = pc.points
    assert pts.shape[0] == 3 and np.isfinite(pts).all()
    plain = get_pointcloud(scene, list(range(n)), maps, False, borders=2).points
    assert plain.shape[1] == n * (H - 4) * (W - 4) and pts.shape[1] <= plain.shape[1]
    # every fused point lies on its pixel's ray at the predicted depth: inside the bounding box (+ one voxel)
    assert (np.abs(plain) <= 1.0 + 2.0 / G + 1e-4).all()


def test_forward_pass_with_filtered_rays(torch_cuda):
    """filter_out_rays=True (forward_pass.py:168-179): only pixels with ground truth become rays; the
    others come back as 0 in the depth map, the rest are positive and finite."""
    from raynet_b200.common.generation_parameters import GenerationParameters
    from raynet_b200.forward_pass import get_forward_pass_factory
    from raynet_b200.synth import SyntheticScene, random_features
    V, H, W, G, D, M = 3, 24, 20, 24, 8, 72
    scene = SyntheticScene(V, H, W, (G, G, G), with_images=True)
    rng = np.random.RandomState(3)
    gt = [(rng.rand(H, W) > 0.4).astype(np.float32) * 2.0 for _ in range(V)]
    scene.get_depth_map = lambda i: gt[i]
    feats = random_features(V, H, W, 32, 11, seed=4)

    class Model(object):
        def predict_features(self, scene, views):
            return feats[list(views)]

    gp = GenerationParameters(depth_planes=D, neighbors=V - 1, grid_shape=np.array([G, G, G], np.int32),
                              max_number_of_marched_voxels=M, padding=11, gamma_mrf=0.05)
    fp = get_forward_pass_factory("raynet")(Model(), gp, "sample_in_bbox", scene.image_shape, H * W, filter_out_rays=True)
    maps = list(fp.forward_pass(scene, (0, V, 1)))
    for k, m in enumerate(maps):
        assert m.shape == (H, W)
        assert (m[gt[k] == 0] == 0).all()
        assert np.isfinite(m).all() and (m[gt[k] != 0] > 0).all()
