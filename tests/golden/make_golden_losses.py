"""Generates tests/golden/losses.npz by EXECUTING THE REFERENCE's own loss code
(/root/reference/utils/tpu.py:26-155 tpu_smooth_l1 / tpu_focal, imported unmodified) on torch-CPU float64
tensors through tests/golden/keras_stub.py's `tensorflow` stand-in; the gradients w.r.t. the predictions come
from torch autograd THROUGH the reference's code.  The one TF-internal piece, keras.backend.binary_crossentropy,
is restated in the stub.  Also evaluates the reference's wBiFPNAdd layer
(layers.py:11-39) on seeded inputs.  Build container only.  Run:  python tests/golden/make_golden_losses.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import keras_stub as ks  # noqa: E402

ks.install()
sys.path.insert(0, "/root/reference")
from utils import tpu  # noqa: E402  (reference utils/tpu.py)
import layers as ref_layers  # noqa: E402  (reference layers.py)

rng = np.random.default_rng(77)
B, N, C = 2, 600, 7
out = {}
state = rng.choice([-1.0, 0.0, 1.0], size=(B, N), p=[0.1, 0.8, 0.1])
labels = np.zeros((B, N, C + 1))
cls_idx = rng.integers(0, C, (B, N))
for b in range(B):
    for n in range(N):
        if state[b, n] == 1:
            labels[b, n, cls_idx[b, n]] = 1
labels[:, :, -1] = state
pred = rng.uniform(0, 1, (B, N, C))
pred[0, :5] = [0.0, 0.999, 1e-9, 0.001, 0.5, 0.3, 0.7]           # lower clip edge (p < 1e-7) and p = .5; the upper edge
# 1 - 1e-7 is not representable in float32 (TF computes the loss in float32), so it is not part of an fp64 fixture
reg_t = np.concatenate([rng.normal(0, 1.5, (B, N, 4)), state[..., None]], -1)
reg_p = rng.normal(0, 1.0, (B, N, 4))
out.update(labels=labels, pred=pred, reg_t=reg_t, reg_p=reg_p)
for tag, (alpha, gamma) in {"a": (0.25, 1.5), "b": (0.25, 2.0)}.items():
    p = torch.tensor(pred, dtype=torch.float64, requires_grad=True)
    loss = tpu.tpu_focal(alpha=alpha, gamma=gamma)(torch.tensor(labels), p)
    loss.backward()
    out["focal_%s" % tag] = np.float64(loss.item())
    out["focal_%s_grad" % tag] = p.grad.numpy()
    out["focal_%s_params" % tag] = np.array([alpha, gamma])
for tag, lam in {"a": 1, "b": 0.5}.items():
    p = torch.tensor(reg_p, dtype=torch.float64, requires_grad=True)
    loss = tpu.tpu_smooth_l1(lam)(torch.tensor(reg_t), p)
    loss.backward()
    out["sl1_%s" % tag] = np.float64(loss.item())
    out["sl1_%s_grad" % tag] = p.grad.numpy()
    out["sl1_%s_lambda" % tag] = np.float64(lam)
# no positives at all: normaliser max(1, 0)
lab0 = labels.copy(); lab0[:, :, :] = 0
out["focal_nopos"] = np.float64(tpu.tpu_focal(0.25, 1.5)(torch.tensor(lab0), torch.tensor(pred)).item())
# wBiFPNAdd
for n_in in (2, 3):
    ks.reset()
    w = rng.uniform(-0.3, 1.0, n_in)
    ks.WEIGHTS = lambda key, shape: w
    xs = [rng.standard_normal((2, 5, 5, 8)) for _ in range(n_in)]
    layer = ref_layers.wBiFPNAdd(name="w_bi_fpn_add")
    y = layer.run([torch.tensor(x) for x in xs])
    out["fuse%d_w" % n_in] = w
    out["fuse%d_x" % n_in] = np.stack(xs)
    out["fuse%d_y" % n_in] = y.numpy()
np.savez_compressed(os.path.join(HERE, "losses.npz"), **out)
print({k: (v.shape if getattr(v, "ndim", 0) else float(v)) for k, v in out.items()})
