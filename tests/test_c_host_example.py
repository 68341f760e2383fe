"""The drop-in boundary is a C ABI: include/effdet_b200.h must be plain C (C99, -pedantic), and a host without Python
or torch must be able to drive the plan level with it.  examples/detect_host.c is that host; here it is compiled
with gcc against the header, linked with the shipped library and run: without a GPU the library has no CPU fallback
and must say so (exit code 3 + effdet_last_error()), with a GPU it must print detections (exit code 0)."""
import os
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "efficientdet_b200")


def _compile(tmp_path, *extra, name="detect_host"):
    exe = str(tmp_path / name)
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", name + ".c"), *extra, "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "hdr.c"
    src.write_text('#include "effdet_b200.h"\nint main(void) { return EFFDET_OK; }\n')
    for std in ("-std=c99", "-std=c11"):
        r = subprocess.run(["gcc", std, "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I",
                            os.path.join(ROOT, "include"), str(src)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True)
        assert r.returncode == 0, r.stdout


def test_c_host_links_and_fails_loudly_without_a_gpu(tmp_path):
    from efficientdet_b200 import _lib
    _lib.load()                                    # the library must exist (built by __graft_entry__.build())
    exe = _compile(tmp_path, "-L", LIBDIR, "-leffdet_b200", "-Wl,-rpath," + LIBDIR, "-lm")
    r = subprocess.run([exe, "9"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked run")
    r = subprocess.run([exe, "0", "128", "1", "4"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 3, (r.stdout, r.stderr)
    assert "effdet_plan_create failed" in r.stderr and len(r.stderr.strip().splitlines()) == 1


def test_c_training_host_links_and_reports_errors(tmp_path):
    """examples/train_host.c (compiled plans: effdet_replay_load / region / step from C, CUDA runtime for the copies)
    compiles as pedantic C99, links, and reports a missing or malformed plan file through effdet_last_error()."""
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("CUDA toolkit headers not found")
    exe = _compile(tmp_path, "-isystem", os.path.join(cuda, "include"), "-L", LIBDIR, "-leffdet_b200",
                   "-Wl,-rpath," + LIBDIR, "-L", os.path.join(cuda, "lib64"), "-lcudart",
                   "-Wl,-rpath," + os.path.join(cuda, "lib64"), name="train_host")
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    r = subprocess.run([exe, str(tmp_path / "missing.efd")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 3 and "cannot open" in r.stderr
    bad = tmp_path / "bad.efd"
    bad.write_bytes(b"NOTAPLAN" + bytes(64))
    r = subprocess.run([exe, str(bad)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 3 and "effdet_replay_load failed (-1)" in r.stderr


@pytest.mark.gpu
def test_c_host_detects_on_the_gpu(tmp_path):
    """The same program on a B200: plan create -> weight manifest -> bind -> detect, no Python in the process."""
    exe = _compile(tmp_path, "-L", LIBDIR, "-leffdet_b200", "-Wl,-rpath," + LIBDIR, "-lm")
    r = subprocess.run([exe, "0", "256", "2", "6"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout, r.stderr)
    lines = r.stdout.strip().splitlines()
    assert lines[0].startswith("EfficientDet-D0 256x256 batch 2, 6 classes: 466 weights, 12276 anchors")
    assert len(lines) == 3 and all("detections" in l for l in lines[1:])
    assert int(lines[1].split()[2]) > 0                 # scores start at the 0.01 prior, threshold 0.005
