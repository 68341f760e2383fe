"""Pins oracle/losses.py and oracle/graph.fuse on fixtures produced by EXECUTING the reference's own
utils/tpu.py (tpu_focal, tpu_smooth_l1) and layers.py (wBiFPNAdd) -- tests/golden/make_golden_losses.py.  CPU."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _z():
    return np.load(os.path.join(HERE, "golden", "losses.npz"))


def test_focal_matches_reference_code():
    from oracle import losses
    z = _z()
    for tag in ("a", "b"):
        alpha, gamma = z["focal_%s_params" % tag]
        p = torch.tensor(z["pred"], requires_grad=True)
        l = losses.focal(torch.tensor(z["labels"]), p, float(alpha), float(gamma))
        l.backward()
        assert abs(l.item() - float(z["focal_%s" % tag])) < 1e-10 * abs(float(z["focal_%s" % tag]))
        g = z["focal_%s_grad" % tag]
        assert np.abs(p.grad.numpy() - g).max() < 1e-10 * np.abs(g).max()
    lab0 = z["labels"].copy()
    lab0[:] = 0
    l0 = losses.focal(torch.tensor(lab0), torch.tensor(z["pred"]), 0.25, 1.5).item()
    assert abs(l0 - float(z["focal_nopos"])) < 1e-10 * float(z["focal_nopos"])


def test_smooth_l1_matches_reference_code():
    from oracle import losses
    z = _z()
    for tag in ("a", "b"):
        p = torch.tensor(z["reg_p"], requires_grad=True)
        l = losses.smooth_l1(torch.tensor(z["reg_t"]), p, float(z["sl1_%s_lambda" % tag]))
        l.backward()
        assert abs(l.item() - float(z["sl1_%s" % tag])) < 1e-12
        assert np.abs(p.grad.numpy() - z["sl1_%s_grad" % tag]).max() < 1e-12


def test_fusion_matches_reference_layer():
    from oracle import graph
    z = _z()
    for n in (2, 3):
        W = {"w/w": z["fuse%d_w" % n]}
        xs = [torch.tensor(x).permute(0, 3, 1, 2) for x in z["fuse%d_x" % n]]
        y = graph.fuse(xs, W, True, "w").permute(0, 2, 3, 1).numpy()
        assert np.abs(y - z["fuse%d_y" % n]).max() < 1e-12


import pytest  # noqa: E402


@pytest.mark.skipif(not os.path.exists("/root/reference/utils/tpu.py"),
                    reason="build container only: executes the reference's utils/tpu.py from where it lies")
@pytest.mark.parametrize("seed", [1, 2])
def test_losses_against_the_reference_executed_live_on_random_problems(seed):
    """Beyond the committed fixture: 16 random problems per seed (batch, anchors, classes, positive / ignore
    fractions incl. none at all, alpha / gamma / lambda) through the reference's own tpu_focal / tpu_smooth_l1 and
    through oracle/losses.py -- values and gradients (tests/golden/check_losses_live.py, a subprocess because the
    Keras stand-in replaces `tensorflow` in sys.modules)."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "check_losses_live.py"), str(seed)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "16 cases agree" in r.stdout, r.stdout[-2000:]
