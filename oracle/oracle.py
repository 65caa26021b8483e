"""ctypes front end of the CPU oracle (oracle/rn_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of rn_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by
anything under raynet_b200/.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rn_oracle.c")
BUILD_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD_DIR, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off -fopenmp rn_oracle.c -> oracle/_build/liboracle.so"""
    os.makedirs(BUILD_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = [
        "gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
        "-fvisibility=hidden", "-Wall", SRC, "-o", LIB, "-lm",
    ]
    subprocess.check_call(cmd)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.rn_oracle_argmax_depth.restype = ctypes.c_float
        _lib.rn_oracle_voxel_traversal.restype = ctypes.c_int
        _lib.rn_oracle_num_threads.restype = ctypes.c_int
    return _lib


def ref_modules():
    """Import the reference's own code compiled by oracle/build_ref.py (oracle/_ref/*.so).

    Returns a dict with the modules that are present: ray_tracing, ref_mrf_np,
    ref_planes_voxels_mapping, ref_mrf_utils."""
    mods = {}
    if not os.path.isdir(REF_DIR):
        return mods
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    for name in ("ray_tracing", "ref_mrf_np", "ref_planes_voxels_mapping", "ref_mrf_utils"):
        try:
            mods[name] = importlib.import_module(name)
        except ImportError:
            pass
    return mods


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads():
    return lib().rn_oracle_num_threads()


def set_threads(n):
    lib().rn_oracle_set_threads(int(n))


# ---------------------------------------------------------------------------- a1
def sample_in_bbox(ray_idxs, H, P_inv, centre, bbox):
    ray_idxs = _i32(ray_idxs)
    N = ray_idxs.shape[0]
    starts = np.zeros((N, 3), np.float32)
    ends = np.zeros((N, 3), np.float32)
    lib().rn_oracle_batch_sample_in_bbox(_p(ray_idxs), ctypes.c_int64(N), int(H), _p(_f32(P_inv).ravel()),
                                         _p(_f32(centre).ravel()), _p(_f32(bbox).ravel()), _p(starts), _p(ends))
    return starts, ends


# ---------------------------------------------------------------------------- a2
def similarity(features, P, starts, ends, D, V, F, H, W, padding):
    features = _f32(features)
    starts, ends = _f32(starts), _f32(ends)
    N = starts.shape[0]
    S = np.zeros((N, D), np.float32)
    lib().rn_oracle_batch_similarity(_p(features), _p(_f32(P).ravel()), _p(starts), _p(ends), ctypes.c_int64(N),
                                     int(D), int(V), int(F), int(H), int(W), int(padding), _p(S))
    return S


def project_pixel(P_view, start, end, k, D, H, W, padding):
    out = np.zeros(2, np.int32)
    lib().rn_oracle_project_pixel(_p(_f32(P_view).ravel()), _p(_f32(start)), _p(_f32(end)), int(k), int(D),
                                  int(H), int(W), int(padding), _p(out))
    return out


# ---------------------------------------------------------------------------- a3
def voxel_traversal(bbox, grid, starts, ends, M):
    starts, ends = _f32(starts).reshape(-1, 3), _f32(ends).reshape(-1, 3)
    N = starts.shape[0]
    idx = np.zeros((N, M, 3), np.int32)
    cnt = np.zeros((N,), np.int32)
    lib().rn_oracle_batch_voxel_traversal(_p(_f32(bbox).ravel()), _p(_i32(grid)), _p(starts), _p(ends),
                                          ctypes.c_int64(N), int(M), _p(idx), _p(cnt))
    return idx, cnt


def voxel_grid(bbox, grid):
    """(Gx, Gy, Gz, 3) voxel centres == get_voxel_grid(...).transpose(1, 2, 3, 0)."""
    grid = _i32(grid)
    out = np.zeros((grid[0], grid[1], grid[2], 3), np.float32)
    lib().rn_oracle_voxel_grid(_p(_f32(bbox).ravel()), _p(grid), _p(out))
    return out


# ---------------------------------------------------------------------------- a4
def planes_voxels_mapping(vgrid, grid, idx, cnt, starts, ends, S, M):
    idx, cnt = _i32(idx), _i32(cnt)
    N = cnt.shape[0]
    S = _f32(S)
    D = S.shape[1]
    S_new = np.zeros((N, M), np.float32)
    lib().rn_oracle_batch_planes_voxels_mapping(_p(_f32(vgrid)), _p(_i32(grid)), _p(idx), _p(cnt),
                                                _p(_f32(starts)), _p(_f32(ends)), _p(S), ctypes.c_int64(N),
                                                int(M), int(D), _p(S_new))
    return S_new


# ---------------------------------------------------------------------------- a5-a8
def _acc_dtype(acc_f64):
    return np.float64 if acc_f64 else np.float32


def bp_iteration(S, idx, cnt, grid, acc_prev, acc_new, msgs, acc_f64=False):
    """One synchronous sweep.  msgs updated IN PLACE, acc_new accumulated IN PLACE."""
    S, idx, cnt = _f32(S), _i32(idx), _i32(cnt)
    N, M = S.shape
    dt = _acc_dtype(acc_f64)
    assert acc_prev.dtype == dt and acc_new.dtype == dt and msgs.dtype == np.float32
    assert acc_new.flags.c_contiguous and msgs.flags.c_contiguous
    lib().rn_oracle_bp_iteration(_p(S), _p(idx), _p(cnt), ctypes.c_int64(N), int(M), _p(_i32(grid)),
                                 _p(np.ascontiguousarray(acc_prev)), _p(acc_new), int(acc_f64), _p(msgs))
    return acc_new, msgs


def belief_propagation(S, idx, cnt, grid, gamma=0.05, bp_iterations=3, acc_f64=False):
    S, idx, cnt = _f32(S), _i32(idx), _i32(cnt)
    N, M = S.shape
    grid = _i32(grid)
    acc = np.zeros(tuple(grid), _acc_dtype(acc_f64))
    msgs = np.zeros((N, M), np.float32)
    lib().rn_oracle_belief_propagation(_p(S), _p(idx), _p(cnt), ctypes.c_int64(N), int(M), _p(grid),
                                       ctypes.c_double(gamma), int(bp_iterations), int(acc_f64), _p(msgs),
                                       _p(acc))
    return acc, msgs


def depth_distribution(S, idx, cnt, grid, acc, msgs, acc_f64=False):
    S, idx, cnt = _f32(S), _i32(idx), _i32(cnt)
    N, M = S.shape
    acc = np.ascontiguousarray(acc, _acc_dtype(acc_f64))
    S_new = np.zeros((N, M), np.float32)
    lib().rn_oracle_depth_distribution(_p(S), _p(idx), _p(cnt), ctypes.c_int64(N), int(M), _p(_i32(grid)),
                                       _p(acc), int(acc_f64), _p(_f32(msgs)), _p(S_new))
    return S_new


def occupancy(acc):
    acc_f64 = acc.dtype == np.float64
    acc = np.ascontiguousarray(acc, _acc_dtype(acc_f64))
    out = np.zeros(acc.shape, np.float32)
    lib().rn_oracle_occupancy(_p(acc), int(acc_f64), ctypes.c_int64(acc.size), _p(out))
    return out


# ---------------------------------------------------------------------------- a9
def argmax_depth(S_new, idx, vgrid, grid, centre):
    S_new, idx = _f32(S_new), _i32(idx)
    N, M = S_new.shape
    depth = np.zeros((N,), np.float32)
    am = np.zeros((N,), np.int32)
    lib().rn_oracle_batch_argmax_depth(_p(S_new), _p(idx), ctypes.c_int64(N), int(M), _p(_f32(vgrid)),
                                       _p(_i32(grid)), _p(_f32(centre).ravel()), _p(depth), _p(am))
    return depth, am


# ---------------------------------------------------------------------------- fused front end
def frontend(ray_idxs, features, P, P_inv, centre, vgrid, bbox, grid, M, D, V, F, H, W, padding,
             want_stages=True):
    ray_idxs = _i32(ray_idxs)
    N = ray_idxs.shape[0]
    idx = np.zeros((N, M, 3), np.int32)
    cnt = np.zeros((N,), np.int32)
    S_vox = np.zeros((N, M), np.float32)
    starts = np.zeros((N, 3), np.float32) if want_stages else None
    ends = np.zeros((N, 3), np.float32) if want_stages else None
    S = np.zeros((N, D), np.float32) if want_stages else None
    lib().rn_oracle_frontend(_p(ray_idxs), ctypes.c_int64(N), _p(_f32(features)), _p(_f32(P).ravel()),
                             _p(_f32(P_inv).ravel()), _p(_f32(centre).ravel()), _p(_f32(vgrid)),
                             _p(_f32(bbox).ravel()), _p(_i32(grid)), int(M), int(D), int(V), int(F), int(H),
                             int(W), int(padding), _p(idx), _p(cnt), _p(S_vox), _p(starts), _p(ends), _p(S))
    return dict(idx=idx, cnt=cnt, S_vox=S_vox, starts=starts, ends=ends, S=S)
