"""ctypes binding of the C-ABI library (include/raynet_b200.h).

There is NO fallback: if libraynet_b200.so is missing or a call fails, this raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libraynet_b200.so")

RN_OK = 0
RN_ERR_SHAPE = 10001
RN_ERR_UNSUPPORTED = 10002
RN_ERR_CUDA = 10003


class RnParams(ctypes.Structure):
    """Mirror of `struct RnParams` (include/raynet_b200.h)."""
    _fields_ = [
        ("max_voxels", ctypes.c_int32),
        ("depth_planes", ctypes.c_int32),
        ("n_views", ctypes.c_int32),
        ("feat_dim", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("width", ctypes.c_int32),
        ("padding", ctypes.c_int32),
        ("grid", ctypes.c_int32 * 3),
        ("bbox", ctypes.c_float * 6),
    ]


def make_params(M=0, D=0, N=0, F=0, H=0, W=0, padding=0, bbox=None, grid_shape=None):
    p = RnParams()
    p.max_voxels, p.depth_planes, p.n_views, p.feat_dim = int(M), int(D), int(N), int(F)
    p.height, p.width, p.padding = int(H), int(W), int(padding)
    if grid_shape is not None:
        for i in range(3):
            p.grid[i] = int(grid_shape[i])
    if bbox is not None:
        import numpy as np
        b = np.asarray(bbox, dtype=np.float32).ravel()
        assert b.shape[0] == 6
        for i in range(6):
            p.bbox[i] = float(b[i])
    return p


_PTR = ctypes.c_void_p
_I64 = ctypes.c_int64
_I32 = ctypes.c_int32
_PP = ctypes.POINTER(RnParams)

# name -> argtypes (all return int unless noted); must list every symbol of the header
SIGNATURES = {
    "rn_sample_in_bbox": [_PP, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _PTR],
    "rn_sample_points": [_PP, _PTR, _PTR, _PTR, _PTR, _I64, _PTR],
    "rn_similarity": [_PP, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _PTR],
    "rn_mvcnn_forward": [_PP, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _PTR],
    "rn_mvcnn_forward_depth": [_PP, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _PTR],
    "rn_voxel_traversal": [_PP, _PTR, _PTR, _PTR, _PTR, _I64, _PTR],
    "rn_planes_to_voxels": [_PP, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _PTR],
    "rn_bp_iteration": [_PP, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _PTR],
    "rn_depth_estimate": [_PP, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _PTR],
    "rn_occupancy": [_PTR, _PTR, _I64, _PTR],
    "rn_fill_f32": [_PTR, ctypes.c_float, _I64, _PTR],
    "rn_raynet_fp": [_PP] + [_PTR] * 12 + [_I64, _PTR],
    "rn_raynet_de": [_PP] + [_PTR] * 12 + [_I64, _PTR],
    "rn_mvcnn_voxel": [_PP] + [_PTR] * 9 + [_I64, _PTR],
    "rn_mvcnn_voxel_depth": [_PP] + [_PTR] * 10 + [_I64, _PTR],
    "rn_engine_frontend": [_PP, _PTR, _PTR, _PTR, _I32] + [_PTR] * 12 + [_I64, _PTR],
    "rn_engine_trace": [_PP] + [_PTR] * 8 + [_I64, _PTR],
    "rn_engine_similarity": [_PP, _PTR, _PTR, _I32] + [_PTR] * 10 + [_I64, _PTR],
    "rn_conv3x3_bn_relu": [_PTR] * 5 + [_I32] * 5 + [_PTR],
    "rn_fuse_depth_maps": [_PTR] * 6 + [_I32] * 5 + [_PTR, _PTR, _PTR],
    "rn_nn_grid_distances": [_PTR, _I64, _PTR, _PTR, _PTR, ctypes.c_float, _PTR, _I32, _PTR, _PTR],
    "rn_engine_bin_rays": [_PP, _PTR, _I64, _I64, _PTR, _PTR, _PTR],
    "rn_engine_bp_iteration": [_PP] + [_PTR] * 8 + [_I32, _I32, _I64, _PTR],
    "rn_engine_depth": [_PP] + [_PTR] * 8 + [_I32, _PTR, _PTR, _I64, _PTR],
    "rn_grid_to_bricks": [_PP, _PTR, _PTR, ctypes.c_float, _PTR],
    "rn_bricks_to_grid": [_PP, _PTR, _PTR, _I32, _PTR],
    "rn_engine_expand_indices": [_PP, _PTR, _PTR, _PTR, _PTR, _I64, _PTR],
    "rn_axis_centres": [_PP, _PTR, _PTR, _PTR],
    "rn_add_prior": [_PTR, ctypes.c_float, _I64, _PTR],
    "rn_max_count": [_PTR, _I64, _PTR, _PTR],
    # parity mode (float64 accumulators)
    "rn_fill_f64": [_PTR, ctypes.c_double, _I64, _PTR],
    "rn_occupancy_f64": [_PTR, _PTR, _I64, _PTR],
    "rn_grid_to_bricks_f64": [_PP, _PTR, _PTR, ctypes.c_double, _PTR],
    "rn_bricks_to_grid_f64": [_PP, _PTR, _PTR, _PTR, _PTR],
    "rn_bp_iteration_f64": [_PP, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _PTR],
    "rn_depth_estimate_f64": [_PP, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _PTR],
    "rn_engine_bp_iteration_f64": [_PP] + [_PTR] * 6 + [_I32, _I64, _PTR],
    "rn_engine_depth_f64": [_PP] + [_PTR] * 8 + [_I32, _PTR, _PTR, _I64, _PTR],
    # backward pass (training)
    "rn_bp_sweep_backward": [_PP] + [_PTR] * 11 + [_I64, _I64, _PTR],
    "rn_depth_estimate_backward": [_PP] + [_PTR] * 10 + [_I64, _I64, _PTR],
    "rn_planes_to_voxels_backward": [_PP] + [_PTR] * 7 + [_I32] + [_PTR] * 3 + [_I64, _PTR],
    "rn_clip_renorm_backward": [_PP] + [_PTR] * 4 + [_I64, _PTR],
    "rn_depth_loss": [_PP, _I32] + [_PTR] * 7 + [ctypes.c_float, _I64, _PTR],
    # mapping fused into the first sweep
    "rn_engine_plane_scores": [_PP, _PTR, _PTR, _I32] + [_PTR] * 4 + [_I64, _PTR],
    "rn_engine_plane_scores_passes": [_PP, _PTR, _PTR, _I32] + [_PTR] * 4 + [_I64, _I32, _PTR],
    "rn_engine_map_planes": [_PP] + [_PTR] * 9 + [_I64, _PTR],
    "rn_engine_first_sweep_mapped": [_PP] + [_PTR] * 14 + [_I64, _PTR],
    # MV-CNN on the tensor cores
    "rn_conv3x3_bn_relu_split": [_PTR] * 6 + [_I32] * 5 + [_PTR],
    "rn_conv3x3_bn_relu_tc": [_PTR] * 7 + [_I32] * 4 + [_PTR],
    # exchange step over NVLink peer memory
    "rn_peer_allreduce_f32": [_PTR, _PTR, _PTR, _PTR, _I32, _I32, _I32, ctypes.c_uint32, ctypes.c_float, _I64, _PTR],
    "rn_peer_allreduce_mc_f32": [_PTR, _PTR, _PTR, _PTR, _I32, _I32, _I32, ctypes.c_uint32, ctypes.c_float, _I64, _PTR],
}
OTHER_SYMBOLS = ["rn_last_error", "rn_abi_version", "rn_device_info", "rn_code_stride", "rn_row_stride", "rn_num_classes",
                 "rn_brick_elems", "rn_backward_scratch_bytes"]

_lib = None


class RayNetB200Error(RuntimeError):
    pass


def load():
    """Load the shared library; raises if it was not built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RayNetB200Error(
            "raynet_b200: %s is missing -- build it with `python -m raynet_b200.build` "
            "(there is no CPU/PyTorch fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    lib.rn_last_error.restype = ctypes.c_char_p
    lib.rn_last_error.argtypes = []
    lib.rn_abi_version.restype = ctypes.c_int
    lib.rn_code_stride.restype = ctypes.c_int64
    lib.rn_code_stride.argtypes = [ctypes.c_int32]
    lib.rn_device_info.argtypes = [ctypes.POINTER(ctypes.c_int)] * 3
    lib.rn_device_info.restype = ctypes.c_int
    lib.rn_row_stride.restype = ctypes.c_int64
    lib.rn_row_stride.argtypes = [ctypes.c_int32]
    lib.rn_num_classes.restype = ctypes.c_int
    lib.rn_num_classes.argtypes = []
    lib.rn_brick_elems.restype = ctypes.c_int64
    lib.rn_brick_elems.argtypes = [_PP]
    lib.rn_backward_scratch_bytes.restype = ctypes.c_int64
    lib.rn_backward_scratch_bytes.argtypes = [_PP, _I64]
    _lib = lib
    return lib


def check(rc):
    """Map a C status to the exception type the reference raises at that point."""
    if rc == RN_OK:
        return
    msg = load().rn_last_error().decode("utf-8", "replace")
    if rc == RN_ERR_SHAPE:
        raise AssertionError(msg)       # the reference asserts on shapes (raynet_fp.py:290-301)
    if rc == RN_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RayNetB200Error("raynet_b200 error %d: %s" % (rc, msg))


def call(name, *args):
    check(getattr(load(), name)(*args))


def code_stride(M):
    return int(load().rn_code_stride(int(M)))


def row_stride(M):
    return int(load().rn_row_stride(int(M)))


def num_classes():
    return int(load().rn_num_classes())


def brick_elems(params):
    n = int(load().rn_brick_elems(ctypes.byref(params)))
    if n < 0:
        check(RN_ERR_SHAPE)
    return n
