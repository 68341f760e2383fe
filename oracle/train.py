"""Oracle: one training step of the reference graph on torch-CPU autograd (TEST INFRASTRUCTURE --
see oracle/__init__.py).  PARITY UNPINNED (TensorFlow cannot run here; no reference fixtures).

Restates what `model.fit` does per step for train_tpu.py:249-346 with --freeze-backbone:
forward with BatchNorm batch statistics in the (trainable) BiFPN and inference-mode BN in the
frozen backbone (TF2 semantics of trainable=False, model.py:57-58), losses utils/tpu.py,
gradients by autograd, Keras SGD (train_tpu.py:268-269), BN moving-average update with
momentum .997 (model.py:42-45; unbiased variance as TF's fused BN does).
"""
import numpy as np
import torch

from . import graph, losses


def trainable_keys(W, freeze_backbone=True, freeze_bn=False):
    keys = []
    for k in W:
        if k.endswith(("moving_mean", "moving_variance")) or k.startswith("boxes/"):
            continue
        is_neck_or_head = k.startswith(("BiFPN_", "w_bi_fpn_add", "box_head/", "class_head/"))
        if freeze_backbone and not is_neck_or_head:
            continue
        if freeze_bn and k.endswith(("/gamma", "/beta")):
            continue
        keys.append(k)
    return keys


def loss_and_grads(W, images, reg_t, lab_t, phi, num_classes, weighted=False, freeze_bn=False,
                   alpha=0.25, gamma=1.5, dtype=torch.float64, freeze_backbone=True, drop_scale=None,
                   force=None):
    """-> (focal, smooth_l1, grads dict, bn batch stats dict name -> (mean, unbiased var)).
    drop_scale: {block prefix: (B,) keep/(1-rate)} = the FixedDropout draw of this step
    (efficientnet.py:300-304), or None for drop_connect_rate=0."""
    keys = trainable_keys(W, freeze_backbone, freeze_bn)
    Wt = {k: torch.tensor(np.asarray(v), dtype=dtype) for k, v in W.items()}
    for k in keys:
        Wt[k].requires_grad_(True)
    stats = {}
    reg, cls = graph.forward(Wt, images, phi, num_classes, weighted, dtype=dtype,
                             bn_train_bifpn=not freeze_bn,
                             bn_train_backbone=(not freeze_backbone) and (not freeze_bn), stats=stats,
                             drop_scale=drop_scale, force=force)
    fl = losses.focal(torch.as_tensor(lab_t).to(dtype), cls, alpha, gamma)
    sl = losses.smooth_l1(torch.as_tensor(reg_t).to(dtype), reg)
    (fl + sl).backward()
    grads = {k: Wt[k].grad.detach().numpy() for k in keys if Wt[k].grad is not None}
    return float(fl.detach()), float(sl.detach()), grads, {k: (m.numpy(), v.numpy()) for k, (m, v) in stats.items()}


def sgd_step(W, grads, velocity, lr=0.01, decay=4e-5, momentum=0.9, iteration=0):
    lr_t = lr / (1.0 + decay * iteration)
    for k, g in grads.items():
        v = velocity.setdefault(k, np.zeros_like(W[k], dtype=np.float64))
        v *= momentum
        v -= lr_t * g
        W[k] = (W[k].astype(np.float64) + v)
    return W
