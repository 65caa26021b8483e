"""Backward pass through the unrolled ray-potential BP -- SURVEY.md 8(f) row 3.

The reference trains end to end through a TensorFlow graph
(raynet/tf_implementations/forward_backward_pass.py:128-248, raynet/mrf/mrf_tf.py:60-271):

    scores -> softmax -> S [N, D] -> plane->voxel interpolation -> S_voxel_space [N, M]
           -> clip_and_renorm -> I unrolled BP sweeps -> depth_estimate -> S_mrf [N, M] -> loss

Here every stage is a `torch.autograd.Function` whose forward AND backward are hand-written kernels
behind the C-ABI (csrc/rn_backward.cuh for the adjoints); PyTorch only carries the tensors and
chains the Functions.  The forward sweeps are the ordinary inference kernels (rn_bp_iteration) with the
inputs of every sweep -- messages and accumulator -- kept as checkpoints; the backward walks them in
reverse.  All tensors are CUDA float32 / int32 in the reference's layouts (voxel lists [N, M, 3],
rows [N, M], accumulators [Gx, Gy, Gz]).

    loss, S_mrf = forward_backward_pass(scores, voxel_grid, idx, cnt, S_target, starts, ends,
                                        camera_centers, grid_shape, gamma, bp_iterations, loss="squared_emd")
    loss.backward()          # scores.grad (and gamma.grad when gamma is a tensor that requires grad)
"""
import numpy as np
import torch

from . import _lib
from .cuda_implementations.utils import current_stream_ptr

LOSS_KINDS = {"emd": 0, "squared_emd": 1, "expected_squared_error": 2}
SCRATCH_BYTES = 1 << 30         # float64 scratch of the adjoint kernels; rays are processed in chunks that fit


def _p(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "training tensors must be contiguous CUDA tensors"
    return t.data_ptr()


def _scratch(params, n_rays, device):
    need = int(_lib.load().rn_backward_scratch_bytes(params, int(n_rays)))
    nbytes = max(int(_lib.load().rn_backward_scratch_bytes(params, 1)), min(need, SCRATCH_BYTES))
    return torch.empty((nbytes // 8,), dtype=torch.float64, device=device), nbytes


class PlanesToVoxels(torch.autograd.Function):
    """S_voxel_space = normalised linear interpolation of softmax(scores) at the ray's voxel centres
    (planes_voxels_mapping.cu:6-92; forward_backward_pass.py:76-125, 176-203).  Gradient w.r.t. scores."""

    @staticmethod
    def forward(ctx, scores, voxel_grid, idx, cnt, starts, ends, grid_shape):
        N, D = scores.shape
        M = idx.shape[1]
        params = _lib.make_params(M=M, D=D, grid_shape=grid_shape)
        S = torch.softmax(scores, dim=1).contiguous()        # K.softmax(S), forward_backward_pass.py:184
        S_vox = torch.zeros((N, M), dtype=torch.float32, device=scores.device)
        _lib.call("rn_planes_to_voxels", params, _p(voxel_grid), _p(idx), _p(cnt), _p(starts), _p(ends), _p(S),
                  _p(S_vox), N, current_stream_ptr())
        ctx.params = params
        ctx.save_for_backward(S, voxel_grid, idx, cnt, starts, ends)
        return S_vox

    @staticmethod
    def backward(ctx, g_S_vox):
        S, voxel_grid, idx, cnt, starts, ends = ctx.saved_tensors
        N = S.shape[0]
        g_S = torch.empty_like(S)
        g_scores = torch.empty_like(S)
        _lib.call("rn_planes_to_voxels_backward", ctx.params, _p(voxel_grid), _p(idx), _p(cnt), _p(starts), _p(ends),
                  _p(S), _p(g_S_vox.contiguous()), 1, None, _p(g_S), _p(g_scores), N, current_stream_ptr())
        return g_scores, None, None, None, None, None, None


class UnrolledBP(torch.autograd.Function):
    """S_mrf = depth_estimate(BP^I(clip_and_renorm(S_voxel_space))) (mrf_tf.py:176-271) with gradients
    w.r.t. S_voxel_space and the occupancy prior gamma."""

    @staticmethod
    def forward(ctx, S_vox, idx, cnt, gamma, grid_shape, bp_iterations):
        N, M = S_vox.shape
        grid_shape = tuple(int(g) for g in grid_shape)
        params = _lib.make_params(M=M, grid_shape=grid_shape)
        dev = S_vox.device
        st = current_stream_ptr()
        g = gamma.detach().to(torch.float64) if isinstance(gamma, torch.Tensor) else torch.tensor(float(gamma), dtype=torch.float64)
        prior = float(torch.log(g) - torch.log(1 - g))
        S_vox = S_vox.contiguous()
        msgs = torch.zeros((N, M), dtype=torch.float32, device=dev)
        acc = torch.full(grid_shape, prior, dtype=torch.float32, device=dev)
        checkpoints = []                      # (acc read by sweep t, messages read by sweep t or None)
        for it in range(int(bp_iterations)):
            checkpoints.append((acc, msgs.clone() if it > 0 else None))
            acc_out = torch.full(grid_shape, prior, dtype=torch.float32, device=dev)
            _lib.call("rn_bp_iteration", params, _p(S_vox), _p(idx), _p(cnt), _p(acc), _p(msgs), _p(acc_out), N, st)
            acc = acc_out
        S_mrf = torch.zeros((N, M), dtype=torch.float32, device=dev)
        _lib.call("rn_depth_estimate", params, _p(S_vox), _p(idx), _p(cnt), _p(acc), _p(msgs), _p(S_mrf), N, st)
        ctx.params, ctx.grid_shape, ctx.checkpoints = params, grid_shape, checkpoints
        ctx.gamma_value = float(g)
        ctx.gamma_is_tensor = isinstance(gamma, torch.Tensor)
        ctx.save_for_backward(S_vox, idx, cnt, acc, msgs)
        return S_mrf

    @staticmethod
    def backward(ctx, g_S_mrf):
        S_vox, idx, cnt, acc_final, msgs_final = ctx.saved_tensors
        N, M = S_vox.shape
        dev = S_vox.device
        st = current_stream_ptr()
        params = ctx.params
        scratch, nbytes = _scratch(params, N, dev)
        g_s = torch.zeros((N, M), dtype=torch.float32, device=dev)          # gradient w.r.t. S_norm, summed over all uses
        g_msg = torch.empty((N, M), dtype=torch.float32, device=dev)
        g_acc = torch.zeros(ctx.grid_shape, dtype=torch.float32, device=dev)
        _lib.call("rn_depth_estimate_backward", params, _p(S_vox), _p(idx), _p(cnt), _p(acc_final), _p(msgs_final),
                  _p(g_S_mrf.contiguous()), _p(g_s), _p(g_msg), _p(g_acc), _p(scratch), nbytes, N, st)
        g_prior = g_acc.sum(dtype=torch.float64)          # every accumulator is prior + sum of messages
        for (acc_in, msg_in) in reversed(ctx.checkpoints):
            g_acc_in = torch.zeros(ctx.grid_shape, dtype=torch.float32, device=dev)
            _lib.call("rn_bp_sweep_backward", params, _p(S_vox), _p(idx), _p(cnt), _p(acc_in), _p(msg_in), _p(g_msg),
                      _p(g_acc), _p(g_s), _p(g_msg), _p(g_acc_in), _p(scratch), nbytes, N, st)
            g_acc = g_acc_in
            g_prior = g_prior + g_acc.sum(dtype=torch.float64)
        g_S = torch.empty((N, M), dtype=torch.float32, device=dev)
        _lib.call("rn_clip_renorm_backward", params, _p(S_vox), _p(cnt), _p(g_s), _p(g_S), N, st)
        g_gamma = None
        if ctx.gamma_is_tensor:
            gv = ctx.gamma_value
            g_gamma = (g_prior * (1.0 / gv + 1.0 / (1.0 - gv))).to(torch.float32)
        return g_S, None, None, g_gamma, None, None


class DepthLoss(torch.autograd.Function):
    """tf_implementations/loss_functions.py:4-35, averaged over the rays (K.mean in forward_backward_pass.py:234-239)."""

    @staticmethod
    def forward(ctx, S_pred, S_target, kind, idx, voxel_grid, camera_centers, grid_shape):
        N, M = S_pred.shape
        params = _lib.make_params(M=M, grid_shape=grid_shape)
        per_ray = torch.empty((N,), dtype=torch.float32, device=S_pred.device)
        g = torch.empty((N, M), dtype=torch.float32, device=S_pred.device)
        _lib.call("rn_depth_loss", params, int(kind), _p(S_target.contiguous()), _p(S_pred.contiguous()), _p(idx),
                  _p(voxel_grid), _p(camera_centers), _p(per_ray), _p(g), 1.0 / N, N, current_stream_ptr())
        ctx.save_for_backward(g)
        return per_ray.mean()

    @staticmethod
    def backward(ctx, g_loss):
        (g,) = ctx.saved_tensors
        return g * g_loss, None, None, None, None, None, None


def depth_loss(loss, S_target, S_pred, idx=None, voxel_grid=None, camera_centers=None, grid_shape=None):
    if loss not in LOSS_KINDS:
        raise KeyError(loss)
    if grid_shape is None:
        grid_shape = (1, 1, 1)
    return DepthLoss.apply(S_pred, S_target, LOSS_KINDS[loss], idx, voxel_grid, camera_centers, grid_shape)


def forward_backward_pass(scores, voxel_grid, ray_voxel_indices, ray_voxel_count, S_target, starts, ends,
                          camera_centers, grid_shape, gamma=0.031, bp_iterations=3, loss="squared_emd"):
    """The graph of forward_backward_pass.py:128-248 from the similarity scores on (the CNN and the feature
    gathers in front of it are outside this row): returns (loss, S_mrf); call loss.backward().

    scores [N, D] (requires_grad), voxel_grid [Gx, Gy, Gz, 3], ray_voxel_indices [N, M, 3] int32,
    ray_voxel_count [N] int32, S_target [N, M], starts / ends [N, 3] (ray segment inside the bounding box; the
    reference passes the D sampled points instead), camera_centers [N, 4]; gamma: float or a 0-d tensor."""
    S_vox = PlanesToVoxels.apply(scores, voxel_grid, ray_voxel_indices, ray_voxel_count, starts, ends,
                                 tuple(int(g) for g in grid_shape))
    S_mrf = UnrolledBP.apply(S_vox, ray_voxel_indices, ray_voxel_count, gamma, tuple(int(g) for g in grid_shape),
                             int(bp_iterations))
    return depth_loss(loss, S_target, S_mrf, ray_voxel_indices, voxel_grid, camera_centers, grid_shape), S_mrf
