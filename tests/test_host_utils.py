"""CPU tests of the host utilities either side of the hot path: the per-epoch
learning-rate schedule (utils/lr_schedule.py:5-68) and un-letterboxing of detections
(inference.py:72-85)."""
import math

import numpy as np
import pytest


def test_cosine_decay_with_linear_warmup_values():
    from efficientdet_b200.utils.lr_schedule import get_cosine_decay_with_linear_warmup
    cb = get_cosine_decay_with_linear_warmup(total_epochs=100, learning_rate_max=0.08, warmup_percent=0.05,
                                             alpha=0.001)
    f = cb.schedule
    # warm-up: lr = start + slope * (epoch_index + 1), slope = 0.08 / 5
    assert f(0, None) == pytest.approx(0.016)
    assert f(4, None) == pytest.approx(0.08)
    # cosine part == tf.keras.experimental.CosineDecay(0.08, 95, alpha=0.001)(epoch_number - 5)
    for e in (5, 20, 57, 99, 150):
        step = min(e + 1 - 5, 95)
        want = 0.08 * ((1 - 0.001) * 0.5 * (1 + math.cos(math.pi * step / 95)) + 0.001)
        assert f(e, None) == pytest.approx(want, rel=1e-12)
    assert f(99, None) == pytest.approx(0.08 * 0.001)          # floor = alpha * max
    with pytest.raises(ValueError):
        get_cosine_decay_with_linear_warmup(10, learning_rate_start=1.0, learning_rate_max=0.5)


def test_scheduler_callback_sets_optimizer_lr():
    from efficientdet_b200.optimizers import SGD
    from efficientdet_b200.utils.lr_schedule import get_cosine_decay_with_linear_warmup

    class M:
        optimizer = SGD(lr=0.5, momentum=0.9)
    cb = get_cosine_decay_with_linear_warmup(total_epochs=10, learning_rate_max=0.08, warmup_percent=0.2)
    cb.set_model(M)
    cb.on_epoch_begin(0)
    assert M.optimizer.lr == pytest.approx(0.04)
    cb.on_epoch_begin(1)
    assert M.optimizer.lr == pytest.approx(0.08)


def test_unletterbox_boxes_matches_reference_arithmetic():
    from efficientdet_b200.utils.postprocess import select_detections, unletterbox_boxes
    rng = np.random.default_rng(0)
    boxes = rng.uniform(-20, 540, (1, 300, 4)).astype(np.float32)
    scale, off_h, off_w, h, w = 0.8, 64, 0, 480, 640
    got = unletterbox_boxes(boxes, scale, off_h, off_w, h, w)
    want = boxes.copy()                                    # inference.py:72-85, line by line
    want[:, :, [0, 2]] = want[:, :, [0, 2]] - off_w
    want[:, :, [1, 3]] = want[:, :, [1, 3]] - off_h
    want /= scale
    want[:, :, 0] = np.clip(want[:, :, 0], 0, w - 1)
    want[:, :, 2] = np.clip(want[:, :, 2], 0, w - 1)
    want[:, :, 1] = np.clip(want[:, :, 1], 0, h - 1)
    want[:, :, 3] = np.clip(want[:, :, 3], 0, h - 1)
    assert np.array_equal(got, want)
    assert boxes is not got and not np.array_equal(boxes, got)     # the input is left untouched
    s = np.array([0.9, 0.2, 0.6], np.float32)
    b, sc, lb = select_detections(got[0, :3], s, np.array([1, 2, 3]), 0.5)
    assert sc.tolist() == pytest.approx([0.9, 0.6]) and lb.tolist() == [1, 3] and b.shape == (2, 4)


def test_lr_schedule_matches_executed_reference():
    """tests/golden/lr_schedule.json: the reference's own get_cosine_decay_with_linear_warmup executed unmodified
    (tests/golden/make_golden_lr.py; only tf.keras.experimental.CosineDecay is restated): every epoch of five
    configurations, incl. the epochs around the warm-up / cosine switch and past the end."""
    import json
    import os
    from efficientdet_b200.utils.lr_schedule import get_cosine_decay_with_linear_warmup
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lr_schedule.json")))
    assert len(cases) == 5
    for c in cases:
        f = get_cosine_decay_with_linear_warmup(**c["kwargs"]).schedule
        got = [f(e, None) for e in range(len(c["lr"]))]
        assert got == pytest.approx(c["lr"], rel=1e-12, abs=1e-15), c["kwargs"]
