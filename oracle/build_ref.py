"""Build the REAL reference into binaries under oracle/_ref/ (test infrastructure only).

This script is the committed recipe the task asks for: it compiles the reference's own
sources *where they lie* under /root/reference (never copied into this repo) into
shared objects that live only in the git-ignored directory ``oracle/_ref/``:

  ray_tracing.*.so              <- raynet/ray_marching/ray_tracing.pyx   (Cython DDA; bit-exact voxel-list oracle)
  ref_mrf_np.*.so               <- raynet/mrf/mrf_np.py                  (NumPy ray-potential BP + depth estimate)
  ref_planes_voxels_mapping.*.so<- raynet/planes_voxels_mapping/planes_voxels_mapping.py (NumPy plane->voxel li / li_2)
  ref_mrf_utils.*.so            <- raynet/mrf/mrf_utils.py               (arg-max -> depth map export)

The three ``.py`` files are Python-2 sources; they are fed to Cython through an
*in-memory* two-regex shim (``print x`` -> ``print(x)``, ``xrange`` -> ``range``) plus
dropping one unused package-relative import. Nothing else is touched, so the compiled
modules execute the reference's own statements. Intermediate files go to a temp dir.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may load these binaries. The product path never does.

Run:  python oracle/build_ref.py  [--reference /root/reference]
It is a no-op (exit 0, message) when the reference tree is absent (e.g. on the GPU box,
which uses the prebuilt files shipped inside oracle/_ref/).
"""
import argparse
import glob
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

PY2_SOURCES = {
    # module name in oracle/_ref  ->  path under the reference tree
    "ref_mrf_np": "raynet/mrf/mrf_np.py",
    "ref_planes_voxels_mapping": "raynet/planes_voxels_mapping/planes_voxels_mapping.py",
    "ref_mrf_utils": "raynet/mrf/mrf_utils.py",
}
PYX_SOURCES = {
    "ray_tracing": "raynet/ray_marching/ray_tracing.pyx",
}


def py2_shim(src):
    """The whole shim: two regexes and one dropped import line."""
    src = re.sub(r"^(\s*)print (.*)$", r"\1print(\2)", src, flags=re.M)
    src = src.replace("xrange(", "range(")
    src = re.sub(r"^from \.\.utils\.generic_utils import .*$", "", src, flags=re.M)
    return src


def _compile_ext(c_file, modname, workdir):
    ext_suffix = sysconfig.get_config_var("EXT_SUFFIX")
    inc = sysconfig.get_paths()["include"]
    import numpy as np
    out = os.path.join(OUT, modname + ext_suffix)
    cmd = [
        "gcc", "-O2", "-fPIC", "-shared", "-fno-strict-aliasing", "-ffp-contract=off",
        "-I", inc, "-I", np.get_include(), c_file, "-o", out, "-lm",
    ]
    subprocess.check_call(cmd, cwd=workdir)
    return out


def build(reference="/root/reference", force=False):
    if not os.path.isdir(os.path.join(reference, "raynet")):
        print("[build_ref] reference tree %s absent: using prebuilt oracle/_ref/ as is" % reference)
        return False
    os.makedirs(OUT, exist_ok=True)
    ext_suffix = sysconfig.get_config_var("EXT_SUFFIX")
    wanted = list(PY2_SOURCES) + list(PYX_SOURCES)
    if not force and all(os.path.exists(os.path.join(OUT, m + ext_suffix)) for m in wanted):
        return True
    workdir = tempfile.mkdtemp(prefix="rn_ref_build_")
    try:
        from Cython.Build import cythonize  # noqa: F401  (checks availability)
        for mod, rel in PYX_SOURCES.items():
            src = os.path.join(reference, rel)
            dst = os.path.join(workdir, mod + ".pyx")
            shutil.copyfile(src, dst)  # temp dir only; never the repo
            subprocess.check_call([sys.executable, "-m", "cython", "-3", dst], cwd=workdir)
            _compile_ext(os.path.join(workdir, mod + ".c"), mod, workdir)
        for mod, rel in PY2_SOURCES.items():
            with open(os.path.join(reference, rel)) as f:
                src = py2_shim(f.read())
            dst = os.path.join(workdir, mod + ".py")
            with open(dst, "w") as f:
                f.write(src)
            subprocess.check_call([sys.executable, "-m", "cython", "-3", dst], cwd=workdir)
            _compile_ext(os.path.join(workdir, mod + ".c"), mod, workdir)
    finally:
        shutil.rmtree(workdir, ignore_errors=True)
    print("[build_ref] built:", sorted(os.path.basename(p) for p in glob.glob(os.path.join(OUT, "*.so"))))
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    build(a.reference, a.force)
