"""Build recipe for oracle/_ref/  (TEST INFRASTRUCTURE -- see oracle/__init__.py).

The only compiled code in the reference is utils/compute_overlap.pyx (Cython).
This script cythonizes it FROM WHERE IT LIES under /root/reference (the shipped
utils/compute_overlap.c targets CPython 3.6/3.7 and cannot be used) and
compiles it with gcc; every output goes to oracle/_ref/ (git-ignored, but it
travels to the GPU box).  No reference source is copied into the repo.

The rest of the reference path is TensorFlow-Keras Python: TensorFlow is not
installed and cannot be installed offline, so it is "unbuildable" here
(DESIGN.md records this).

Also compiles the oracle's own C restatement (oracle/overlap.c) into
oracle/_build/liboracle_overlap.so.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
REF_OUT = os.path.join(HERE, "_ref")
OWN_OUT = os.path.join(HERE, "_build")


def _run(cmd):
    subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def build_own(force=False):
    os.makedirs(OWN_OUT, exist_ok=True)
    src = os.path.join(HERE, "overlap.c")
    out = os.path.join(OWN_OUT, "liboracle_overlap.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        # -ffp-contract=off: the reference's Cython module is built without FMA
        # contraction (plain x86-64), keep the double arithmetic bit-identical.
        _run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", out, "-lm"])
    return out


def build_ref(force=False):
    pyx = os.path.join(REF, "utils", "compute_overlap.pyx")
    if not os.path.exists(pyx):
        return None          # e.g. on the GPU box: use the prebuilt file if present
    import numpy
    os.makedirs(REF_OUT, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(REF_OUT, "compute_overlap" + ext)
    if force or not os.path.exists(out):
        c_file = os.path.join(REF_OUT, "compute_overlap.c")
        _run([sys.executable, "-m", "cython", "-3", "-o", c_file, pyx])
        _run(["gcc", "-O2", "-shared", "-fPIC", "-w",
              "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
              c_file, "-o", out])
    return out


def load_ref_compute_overlap():
    """Returns the reference's own compute_overlap (compiled) or None."""
    import importlib.util
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    path = os.path.join(REF_OUT, "compute_overlap" + ext)
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("compute_overlap", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.compute_overlap


if __name__ == "__main__":
    print("own:", build_own(force=True))
    print("ref:", build_ref(force=True))
