"""TEST INFRASTRUCTURE ONLY (never imported by raynet_b200/): a float64 torch restatement of the reference's
differentiable graph, so that torch.autograd yields the reference gradients the backward kernels
(raynet_b200/csrc/rn_backward.cuh) are checked against.

Follows, statement for statement,
  * raynet/mrf/mrf_tf.py:6-15    clip_and_renorm
  * raynet/mrf/mrf_tf.py:18-57   extract_occupancy_to_ray_pos
  * raynet/mrf/mrf_tf.py:60-143  single_ray_belief_propagation
  * raynet/mrf/mrf_tf.py:146-173 single_ray_depth_estimate
  * raynet/mrf/mrf_tf.py:176-271 belief_propagation / depth_estimate (synchronous sweeps, acc = sum + prior)
  * raynet/tf_implementations/loss_functions.py:4-35 emd / squared_emd / expected_squared_error
  * raynet/tf_implementations/forward_backward_pass.py:76-125 single_ray_depth_to_voxels_map_li, in the flavour
    of cuda_implementations/planes_voxels_mapping.cu:36-92 that the kernels implement (t clamped to
    [1e-4, 1 - 1e-4], the two planes around t; identical to the TF top-2 rule for t inside [0, 1]).
TensorFlow is not installed here, so this file is pinned differently: its forward pass must reproduce
oracle/rn_oracle.c (itself pinned against executions of the reference's mrf_np) -- tests/test_oracle_pinning.py --
and its gradients are checked against float64 finite differences of itself (torch.autograd.gradcheck).
Plain Python loops over rays: small cases only.
"""
import numpy as np
import torch

DT = torch.float64


def clip_and_renorm(S, cnt, eps=1e-5):
    N, M = S.shape
    clipped = torch.clamp(S, eps, 1 - eps)
    S_sum = clipped.sum(dim=1, keepdim=True) - (M - cnt.reshape(-1, 1).to(S.dtype)) * eps
    return clipped / S_sum


def occupancy_to_ray(acc_at_voxels, msgs):
    x = acc_at_voxels - msgs
    m = torch.clamp(x, min=0.0)
    t1 = torch.exp(0.0 - m)
    t2 = torch.exp(x - m)
    return torch.clamp(t2 / (t1 + t2), 1e-4, 1 - 1e-4)


def _excl_cumprod(v):
    return torch.cat([torch.ones(1, dtype=v.dtype), torch.cumprod(v, 0)[:-1]])


def _excl_cumsum(v):
    return torch.cat([torch.zeros(1, dtype=v.dtype), torch.cumsum(v, 0)[:-1]])


def single_ray_bp(s, acc_at_voxels, msgs):
    o = occupancy_to_ray(acc_at_voxels, msgs)
    cp = _excl_cumprod(1.0 - o)
    common = cp * s
    new_common = _excl_cumsum(o * common)
    pos = common + new_common
    a = o * common
    t1 = torch.flip(_excl_cumsum(torch.flip(a, [0])), [0])       # reverse exclusive cumsum
    neg = new_common + t1 / (1.0 - o)
    p = pos / (pos + neg)
    return torch.log(p) - torch.log(1.0 - p)


def single_ray_depth(s, acc_at_voxels, msgs):
    o = occupancy_to_ray(acc_at_voxels, msgs)
    P = o * _excl_cumprod(1.0 - o) * s
    return P / P.sum()


def belief_propagation(S_norm, lin, cnt, n_voxels, gamma, bp_iterations):
    """lin: list of int64 tensors (linear voxel index per traversed voxel).  Returns (acc [n_voxels], msgs list)."""
    N = len(lin)
    prior = torch.log(gamma) - torch.log(1.0 - gamma)
    acc = prior * torch.ones(n_voxels, dtype=DT)
    msgs = [torch.zeros(int(c), dtype=DT) for c in cnt]
    for _ in range(bp_iterations):
        new_msgs, parts = [], []
        for r in range(N):
            c = int(cnt[r])
            if c <= 1:                       # mrf_np.py:299-301 (the TF map over an empty / one-voxel slice adds nothing useful)
                new_msgs.append(msgs[r])
                continue
            m = single_ray_bp(S_norm[r, :c], acc[lin[r]], msgs[r])
            new_msgs.append(m)
            parts.append((lin[r], m))
        acc_new = torch.zeros(n_voxels, dtype=DT)
        for (l, m) in parts:
            acc_new = acc_new.index_add(0, l, m)
        acc = acc_new + prior
        msgs = new_msgs
    return acc, msgs


def depth_estimate(S_norm, lin, cnt, acc, msgs, M):
    rows = []
    for r in range(len(lin)):
        c = int(cnt[r])
        if c <= 1:
            rows.append(torch.zeros(M, dtype=DT))
            continue
        P = single_ray_depth(S_norm[r, :c], acc[lin[r]], msgs[r])
        rows.append(torch.cat([P, torch.zeros(M - c, dtype=DT)]))
    return torch.stack(rows)


def planes_to_voxels(S, centres, cnt, starts, ends, M):
    """S [N, D] plane distributions; centres: list of [c, 3] voxel centres per ray -> S_voxel_space [N, M]."""
    N, D = S.shape
    rows = []
    step = 1.0 / (D - 1)
    for r in range(N):
        c = int(cnt[r])
        if c == 0:
            rows.append(torch.zeros(M, dtype=DT))
            continue
        ray = (ends[r] - starts[r]).to(DT)
        t = ((centres[r].to(DT) - starts[r].to(DT)) @ ray) / (ray @ ray)
        t = torch.clamp(t, 1e-4, 1 - 1e-4)
        left = torch.clamp(torch.floor(t / step), max=D - 2).to(torch.int64)
        dl = t - left.to(DT) * step
        dr = (left + 1).to(DT) * step - t
        c1 = 1.0 - dl / (dl + dr)
        c2 = 1.0 - dr / (dl + dr)
        u = c1 * S[r, left] + c2 * S[r, left + 1]
        u = u / u.sum()
        rows.append(torch.cat([u, torch.zeros(M - c, dtype=DT)]))
    return torch.stack(rows)


def emd(y_true, y_pred):
    return torch.abs(torch.cumsum(y_true - y_pred, dim=-1)).mean(dim=-1)


def squared_emd(y_true, y_pred):
    return (torch.cumsum(y_true - y_pred, dim=-1) ** 2).sum(dim=-1)


def expected_squared_error(y_true, y_pred, dists):
    return torch.abs((y_true * dists).sum(-1) - (y_pred * dists).sum(-1))


def forward_graph(scores, idx, cnt, grid, vgrid, starts, ends, centres_cam, S_target, gamma, bp_iterations, loss):
    """The whole graph on numpy / torch inputs (idx int32 [N, M, 3], vgrid float32 [Gx, Gy, Gz, 3]).
    scores: float64 tensor [N, D] (requires_grad as the caller wishes); gamma: float64 0-d tensor.
    Returns (loss, S_mrf, S_voxel_space)."""
    N, M = idx.shape[0], idx.shape[1]
    gx, gy, gz = (int(g) for g in grid)
    lin, centres = [], []
    for r in range(N):
        c = int(cnt[r])
        v = torch.from_numpy(np.asarray(idx[r, :c], np.int64))
        lin.append((v[:, 0] * gy + v[:, 1]) * gz + v[:, 2])
        centres.append(torch.from_numpy(np.asarray(vgrid[idx[r, :c, 0], idx[r, :c, 1], idx[r, :c, 2]], np.float64)))
    cnt_t = torch.from_numpy(np.asarray(cnt, np.int64))
    S = torch.softmax(scores, dim=1)
    S_vox = planes_to_voxels(S, centres, cnt, torch.from_numpy(np.asarray(starts, np.float64)),
                             torch.from_numpy(np.asarray(ends, np.float64)), M)
    S_norm = clip_and_renorm(S_vox, cnt_t)
    acc, msgs = belief_propagation(S_norm, lin, cnt, gx * gy * gz, gamma, bp_iterations)
    S_mrf = depth_estimate(S_norm, lin, cnt, acc, msgs, M)
    y_true = torch.from_numpy(np.asarray(S_target, np.float64))
    if loss == "emd":
        per_ray = emd(y_true, S_mrf)
    elif loss == "squared_emd":
        per_ray = squared_emd(y_true, S_mrf)
    elif loss == "expected_squared_error":
        # gather_nd over ALL M slots: padded slots index voxel (0, 0, 0) (loss_functions.py:22-29)
        vc = torch.from_numpy(np.asarray(vgrid[idx[..., 0], idx[..., 1], idx[..., 2]], np.float64))
        cam = torch.from_numpy(np.asarray(centres_cam, np.float64))[:, None, :3]
        dists = torch.sqrt(((vc - cam) ** 2).sum(-1))
        per_ray = expected_squared_error(y_true, S_mrf, dists)
    else:
        raise KeyError(loss)
    return per_ray.mean(), S_mrf, S_vox
