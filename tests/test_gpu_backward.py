"""SURVEY.md 8(f) row 3 on the GPU: the backward pass through the unrolled ray-potential BP
(raynet_b200/training.py -> csrc/rn_backward.cuh through the C-ABI) against the reference gradients
obtained by torch.autograd from the float64 restatement of the reference's TensorFlow graph
(oracle/bp_autograd.py, pinned against rn_oracle.c in the CPU suite), plus a finite-difference spot check
against rn_oracle.c itself.
"""
import numpy as np
import pytest

from test_oracle_pinning import _training_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "the gpu-marked tests need a CUDA device"
    return torch


def _gpu_run(torch, c, o, scores, target, loss, iters, gamma=0.05):
    from raynet_b200.training import forward_backward_pass
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    z = d(scores.astype(np.float32)).requires_grad_(True)
    g = torch.tensor(gamma, dtype=torch.float32, device="cuda", requires_grad=True)
    cam = np.tile(np.append(c.centre[:3], 1.0).astype(np.float32), (c.N, 1))
    L, S_mrf = forward_backward_pass(z, d(c.vgrid), d(o["idx"]), d(o["cnt"]), d(target), d(o["starts"]), d(o["ends"]),
                                     d(cam), c.grid, gamma=g, bp_iterations=iters, loss=loss)
    L.backward()
    return float(L.detach()), S_mrf.detach().cpu().numpy(), z.grad.cpu().numpy(), float(g.grad)


@pytest.mark.parametrize("loss", ["squared_emd", "emd", "expected_squared_error"])
@pytest.mark.parametrize("iters", [1, 3])
def test_backward_matches_autograd_of_the_reference_graph(torch_cuda, oracle, loss, iters, parity_log):
    torch = torch_cuda
    from oracle import bp_autograd as ag
    c, o, scores, target = _training_case(oracle, n_rays=300, seed=7)
    L, S_mrf, gz, gg = _gpu_run(torch, c, o, scores, target, loss, iters)
    z = torch.from_numpy(scores).requires_grad_(True)
    g = torch.tensor(0.05, dtype=torch.float64, requires_grad=True)
    cam = np.tile(np.append(c.centre[:3], 1.0), (c.N, 1))
    L_ref, S_ref, _ = ag.forward_graph(z, o["idx"], o["cnt"], c.grid, c.vgrid, o["starts"], o["ends"], cam, target, g,
                                       iters, loss)
    L_ref.backward()
    gz_ref, gg_ref = z.grad.numpy(), float(g.grad)
    e_fwd = float(np.abs(S_mrf - S_ref.detach().numpy()).max())
    scale = float(np.abs(gz_ref).max())
    e_gz = float(np.abs(gz - gz_ref).max() / scale)
    e_gg = abs(gg - gg_ref) / max(abs(gg_ref), 1e-12)
    parity_log["backward/%s/%d_sweeps" % (loss, iters)] = {
        "S_mrf_forward": e_fwd, "loss_rel": abs(L - float(L_ref)) / abs(float(L_ref)),
        "grad_scores_rel_to_max": e_gz, "grad_gamma_rel": e_gg, "max_abs_grad_scores": scale}
    assert e_fwd <= 1e-5
    assert abs(L - float(L_ref)) <= 1e-4 * abs(float(L_ref))
    assert e_gz <= 2e-5, e_gz           # measured 4e-7 .. 6e-7 (profiles/r02_parity.json): float32 checkpoints and atomics
    assert e_gg <= 2e-5, (gg, gg_ref)
    # direction: the cosine between the two gradients
    cos = float((gz * gz_ref).sum() / (np.linalg.norm(gz) * np.linalg.norm(gz_ref)))
    assert cos > 1 - 1e-5


def test_backward_finite_differences_of_the_c_oracle(torch_cuda, oracle):
    """Central differences of the C oracle's forward pass (float64 flavour) along a few score directions."""
    torch = torch_cuda
    c, o, scores, target = _training_case(oracle, n_rays=200, seed=11)
    _, _, gz, _ = _gpu_run(torch, c, o, scores, target, "squared_emd", 2)

    def loss_of(sc):
        e = np.exp(sc - sc.max(axis=1, keepdims=True))
        S = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
        S_vox = oracle.planes_voxels_mapping(c.vgrid, c.grid, o["idx"], o["cnt"], o["starts"], o["ends"], S, c.M)
        acc, msgs = oracle.belief_propagation(S_vox, o["idx"], o["cnt"], c.grid, gamma=0.05, bp_iterations=2, acc_f64=True)
        P = oracle.depth_distribution(S_vox, o["idx"], o["cnt"], c.grid, acc, msgs, acc_f64=True).astype(np.float64)
        return float((np.cumsum(target.astype(np.float64) - P, axis=1) ** 2).sum(axis=1).mean())

    rng = np.random.RandomState(0)
    for _ in range(4):
        direction = rng.randn(*scores.shape)
        direction /= np.linalg.norm(direction)
        h = 0.05
        fd = (loss_of(scores + h * direction) - loss_of(scores - h * direction)) / (2 * h)
        an = float((gz.astype(np.float64) * direction).sum())
        assert abs(fd - an) <= 0.03 * max(abs(an), abs(fd)) + 1e-7, (fd, an)


def test_backward_degenerate_rays_and_chunking(torch_cuda, oracle):
    """Rays with count <= 1 get zero gradients; a scratch that holds only a few rays (chunked launches)
    gives the same result as one that holds them all."""
    torch = torch_cuda
    from raynet_b200 import training
    c, o, scores, target = _training_case(oracle, n_rays=150, seed=13)
    o["cnt"][:5] = 0
    o["cnt"][5:9] = 1
    target[:9] = 0
    full = _gpu_run(torch, c, o, scores, target, "squared_emd", 2)
    old = training.SCRATCH_BYTES
    try:
        training.SCRATCH_BYTES = 7 * c.M * 8 * 16          # 16 rays per launch
        small = _gpu_run(torch, c, o, scores, target, "squared_emd", 2)
    finally:
        training.SCRATCH_BYTES = old
    assert np.all(full[2][:9] == 0)
    assert np.abs(full[2] - small[2]).max() <= 1e-6 * np.abs(full[2]).max() + 1e-9
    assert abs(full[0] - small[0]) <= 1e-6 * abs(full[0])
