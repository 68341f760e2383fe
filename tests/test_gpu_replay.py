"""Compiled training plans (plan-level C ABI, training half): a plan exported by plan_export.export_train_plan and
replayed by the C library through ctypes with numpy host buffers only (efficientdet_b200.plan.CReplay) trains exactly
like Trainer.train_on_batch -- same losses, same weights, bit for bit -- for a frozen and for a trained backbone."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batches(B, S, C, kmax, n, seed=3):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        img = rng.standard_normal((B, S, S, 3)).astype(np.float32)
        boxes = np.zeros((B, kmax, 4), np.float64)
        labels = np.zeros((B, kmax), np.int32)
        counts = np.zeros((B,), np.int32)
        for b in range(B):
            k = int(rng.integers(1, kmax + 1))
            wh = rng.uniform(S * 0.15, S * 0.5, (k, 2))
            xy = rng.uniform(0, 1, (k, 2)) * (S - wh)
            boxes[b, :k] = np.concatenate([xy, xy + wh], 1)
            labels[b, :k] = rng.integers(0, C, k)
            counts[b] = k
        out.append((img, boxes, labels, counts))
    return out


@pytest.mark.parametrize("freeze_backbone,dtype", [(True, "bf16"), (False, "bf16"), (True, "fp32")])
def test_compiled_plan_trains_like_the_python_trainer(tmp_path, freeze_backbone, dtype):
    from efficientdet_b200 import _lib
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.optimizers import SGD
    from efficientdet_b200.plan import CReplay
    from efficientdet_b200.plan_export import export_train_plan
    from efficientdet_b200.utils.anchors import anchors_for_shape
    B, S, C, kmax, steps = 2, 128, 4, 6, 3

    def make():
        m = efficientdet(0, num_classes=C, weighted_bifpn=True, image_size=S, dtype=dtype, just_training_model=True,
                         seed=11, drop_connect_rate=0.2)
        if freeze_backbone:
            m.freeze_backbone()
        m.compile(optimizer=SGD(lr=0.02, decay=1e-3, momentum=0.9))
        return m
    data = _batches(B, S, C, kmax, steps)

    # --- the Python trainer (device target assignment + captured graph + SGD), the path bench.py times
    m1 = make()
    tr = m1._trainer
    plan = tr.plan(B, False)
    dev = m1.net.device
    anchors_d = torch.from_numpy(anchors_for_shape((S, S))).to(dev)
    hw = torch.full((B, 2), float(S), dtype=torch.float64, device=dev)
    want_losses = []
    for img, boxes, labels, counts in data:
        plan.tensor(plan.input_images).copy_(torch.from_numpy(img).to(dev))
        tr.targets_into_plan(plan, anchors_d, torch.from_numpy(boxes).to(dev), torch.from_numpy(labels).to(dev),
                             torch.from_numpy(counts).to(dev), hw, kmax)
        tr.run_step(plan)
        want_losses.append(plan.tensor(plan.loss_out).cpu().numpy().copy())
    want_w = m1.net.flat.cpu().numpy().copy()

    # --- the same model exported BEFORE any step, replayed by the C library
    m2 = make()
    path = os.path.join(str(tmp_path), "plan.efd")
    info = export_train_plan(m2, B, path, kmax=kmax)
    meta = json.load(open(path + ".json"))
    assert meta["batch"] == B and meta["kmax"] == kmax and "stem_conv/kernel" in meta["weights"]
    del m2
    rp = CReplay(path)
    assert rp.num_launches == info["ops"] > 100
    got_losses = []
    for i, (img, boxes, labels, counts) in enumerate(data):
        rp.write("images", img)
        rp.write("gt_boxes", boxes)
        rp.write("gt_labels", labels)
        rp.write("gt_counts", counts)
        rp.step(info["lr"] / (1.0 + info["decay"] * i))
        torch.cuda.synchronize()
        got_losses.append(rp.read("losses", np.float32, (8,)))
    got_w = rp.read("weights", np.float32)[:want_w.size]
    rp.close()
    for a, b in zip(got_losses, want_losses):
        assert np.array_equal(a[:2], b[:2]), (got_losses, want_losses)
    assert np.isfinite(want_w).all() and np.array_equal(got_w, want_w)
    assert not np.array_equal(want_w, make().net.flat.cpu().numpy())      # the steps did change the weights
