"""Build the C-ABI shared library (raynet_b200/libraynet_b200.so) with nvcc for sm_100a.

In-tree build: the .so is git-ignored but travels to the GPU box with the snapshot.
-fmad=false: see csrc/rn_common.cuh (bit-exact integer decisions vs the CPU oracle).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "rn_api.cu")
DEPS = [SRC] + [os.path.join(HERE, "csrc", h) for h in ("rn_kernels.cuh", "rn_engine.cuh", "rn_bp4.cuh", "rn_bp4c.cuh", "rn_parity.cuh", "rn_backward.cuh", "rn_peer.cuh", "rn_first.cuh", "rn_cnn_tc.cuh", "rn_simmap3.cuh", "rn_cnn.cuh", "rn_fusion.cuh", "rn_common.cuh")] + [
    os.path.join(ROOT, "include", "raynet_b200.h")]
LIB = os.path.join(HERE, "libraynet_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    subprocess.check_call(cmd)
    return LIB


def build_variant(name, defines):
    """Experiment builds (scripts/gpu_variants.sh): the same library with -D overrides of the tuning macros,
    written to raynet_b200/variants/<name>.so; a run copies one over libraynet_b200.so."""
    out_dir = os.path.join(HERE, "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, name + ".so")
    cmd = [os.environ.get("NVCC", "nvcc")] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out, SRC]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
