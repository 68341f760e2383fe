"""Live check (build container only): the reference's own utils/tpu.py losses (tpu_focal, tpu_smooth_l1, imported
unmodified through tests/golden/keras_stub.py) against oracle/losses.py on random problems -- values and gradients
w.r.t. the predictions (torch autograd THROUGH the reference's code), float64.  tests/test_oracle_losses_golden.py
runs it in a subprocess (the stand-in replaces `tensorflow` in sys.modules).  Exit code 0 = all cases agree."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import keras_stub as ks  # noqa: E402
from oracle import losses  # noqa: E402

ks.install()
sys.path.insert(0, "/root/reference")
from utils import tpu  # noqa: E402  (reference utils/tpu.py)

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
worst = 0.0
for case in range(16):
    B, N, C = int(rng.integers(1, 4)), int(rng.integers(1, 400)), int(rng.integers(1, 12))
    p_pos, p_ign = float(rng.choice([0.0, 0.02, 0.3])), float(rng.choice([0.0, 0.1, 0.5]))
    state = rng.choice([1.0, -1.0, 0.0], size=(B, N), p=[p_pos, p_ign, 1 - p_pos - p_ign])
    labels = np.zeros((B, N, C + 1))
    cls = rng.integers(0, C, (B, N))
    labels[np.arange(B)[:, None], np.arange(N)[None], cls] = (state == 1)
    labels[:, :, -1] = state
    pred = rng.uniform(0, 1, (B, N, C)) ** float(rng.choice([1, 3]))          # includes values near 0
    alpha, gamma = float(rng.choice([0.25, 0.5])), float(rng.choice([1.5, 2.0]))
    reg_t = np.concatenate([rng.normal(0, 2.0, (B, N, 4)), state[..., None]], -1)
    reg_p = rng.normal(0, 1.0, (B, N, 4))
    lam = float(rng.choice([0.5, 1.0]))
    for ref_fn, ora_fn, t, p in (
            (tpu.tpu_focal(alpha=alpha, gamma=gamma), lambda a, b: losses.focal(a, b, alpha, gamma), labels, pred),
            (tpu.tpu_smooth_l1(lam), lambda a, b: losses.smooth_l1(a, b, lam), reg_t, reg_p)):
        out = []
        for fn in (ref_fn, ora_fn):
            x = torch.tensor(p, dtype=torch.float64, requires_grad=True)
            l = fn(torch.tensor(t), x)
            l.backward()
            out.append((l.item(), x.grad.numpy()))
        (lr, gr), (lo, go) = out
        err_l = abs(lr - lo) / max(abs(lr), 1e-30)
        err_g = np.abs(gr - go).max() / max(np.abs(gr).max(), 1e-30)
        worst = max(worst, err_l, err_g)
        if err_l > 1e-10 or err_g > 1e-10:
            print("MISMATCH case", case, dict(B=B, N=N, C=C, alpha=alpha, gamma=gamma, lam=lam), lr, lo, err_l, err_g)
            sys.exit(1)
print("16 cases agree, worst relative difference %.3g" % worst)
