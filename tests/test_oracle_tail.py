"""Pins the oracle's detection tail on the reference's own test vectors
(test_RegressBoxes.py:8-46, test_ClipBoxes.py:14-51, test_FilterDetections.py:8-28)."""
import numpy as np

from oracle import tail


def test_apply_bbox_deltas_reference_vector():
    boxes = np.array([[[0, 0, 1, 1], [0.5, 0.5, 0.6, 0.6]]], "float32")
    deltas = np.array([[[0.1, 0.1, 0.1, 0.1], [-0.2, -0.2, 0.2, 0.2]]], "float32")
    want = np.array([[[0.02, 0.02, 1.02, 1.02], [0.496, 0.496, 0.604, 0.604]]])
    np.testing.assert_array_almost_equal(tail.apply_bbox_deltas(boxes, deltas), want)


def test_clip_boxes_reference_vector():
    boxes = np.array([[[-0.1, 0, 1.3, 1], [10, 0, 210, 300], [-100, -0.5, 0.6, 180]]], np.float32)
    want = np.array([[[0, 0, 1.3, 1], [10, 0, 199, 199], [0, 0, 0.6, 180]]], np.float32)
    np.testing.assert_array_equal(tail.clip_boxes((32, 200, 200, 3), boxes), want)


def test_filter_by_score_and_nms_reference_vector():
    boxes = np.array([[0, 0, 1, 1], [0.5, 0.5, 0.6, 0.6], [0.1, 0.1, 0.6, 0.6]], "float32")
    scores = np.array([.6, .2, .1], "float32")
    labels = np.array([1, 2, 1], "int64")
    got = tail.filter_by_score_and_nms(scores, labels, .12, boxes, 3, .5)
    np.testing.assert_array_equal(got, np.array([[0, 1], [1, 2]]))


def test_filter_detections_pad_and_order():
    rng = np.random.default_rng(0)
    xy = rng.uniform(0, 100, (50, 2)).astype(np.float32)
    boxes = np.concatenate([xy, xy + rng.uniform(5, 30, (50, 2)).astype(np.float32)], 1)
    cls = rng.uniform(0, 1, (50, 3)).astype(np.float32)
    b, s, l = tail.filter_detections(boxes, cls, score_threshold=0.5, max_detections=20)
    assert b.shape == (20, 4) and s.shape == (20,) and l.dtype == np.int32
    k = int((s >= 0).sum())
    assert np.all(np.diff(s[:k]) <= 0)
    assert np.all(s[k:] == -1) and np.all(l[k:] == -1) and np.all(b[k:] == -1)
    # no-NMS path keeps everything above the threshold (up to max_detections)
    b2, s2, l2 = tail.filter_detections(boxes, cls, score_threshold=0.5, max_detections=200,
                                        iou_threshold=0)
    assert int((s2 >= 0).sum()) == int((cls > 0.5).sum())


# ---------------------------------------------------------------- fixtures from the EXECUTED reference
# tests/golden/filter_detections.npz: outputs of the reference's own FilterDetections.py (layer, filter_detections,
# filter_by_score_and_nms) run unmodified over a numpy stand-in for its TensorFlow ops
# (tests/golden/make_golden_filter.py): pins the oracle's thresholding, per-class order, top-k, padding, casts,
# nms=False and class_specific_filter=False handling on the reference's control flow, bit for bit.
import ast
import os

import pytest

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filter_detections.npz")
CASES = ("pad", "topk", "class_cap", "max_class", "no_nms", "no_nms_max_class", "tight_iou", "nothing")


@pytest.mark.parametrize("case", CASES)
def test_filter_detections_layer_matches_executed_reference(case):
    z = np.load(FIX)
    kw = dict(ast.literal_eval(str(z[case + "/kw"])))
    b, s, l = tail.filter_detections_batch(z[case + "/boxes_in"], z[case + "/cls_in"], **kw)
    assert l.dtype == np.int32 and s.dtype == np.float32 and b.dtype == np.float32
    np.testing.assert_array_equal(l, z[case + "/labels"])
    np.testing.assert_array_equal(s, z[case + "/scores"])
    np.testing.assert_array_equal(b, z[case + "/boxes"])


def test_fixture_cases_exercise_what_they_claim():
    z = np.load(FIX)
    assert (z["pad/labels"] == -1).any() and (z["pad/labels"] >= 0).any()             # padding present
    assert (z["topk/labels"] >= 0).all()                                               # top-k cut
    assert (z["nothing/labels"] == -1).all() and (z["nothing/boxes"] == -1).all()
    lab = z["class_cap/labels"][0]
    assert (lab >= 0).all() and len(set(lab.tolist())) == 2
    assert np.all(np.diff(z["topk/scores"], axis=1) <= 0)                              # descending scores


def test_filter_detections_function_and_index_form_match_executed_reference():
    z = np.load(FIX)
    b, s, l = tail.filter_detections(z["fn/boxes_in"], z["fn/cls_in"], score_threshold=0.6, max_detections=35,
                                     iou_threshold=0.4)
    np.testing.assert_array_equal(l, z["fn/labels"])
    np.testing.assert_array_equal(s, z["fn/scores"])
    np.testing.assert_array_equal(b, z["fn/boxes"])
    idx = tail.filter_by_score_and_nms(z["fn/cls_in"][:, 1], z["fsn/labels_in"], 0.5, z["fn/boxes_in"], 20, 0.45)
    np.testing.assert_array_equal(idx, z["fsn/indices"])


def test_decode_and_clip_match_executed_reference():
    """apply_bbox_deltas (RegressBoxes.py:126-164, default and custom mean / std) and ClipBoxes.call
    (ClipBoxes.py:9-24) executed from the reference on 3 x 18 414 boxes: float32 results, bit for bit."""
    z = np.load(FIX)
    a, d = z["dc/anchors"], z["dc/deltas"]
    np.testing.assert_array_equal(tail.apply_bbox_deltas(np.broadcast_to(a, d.shape), d), z["dc/decoded"])
    np.testing.assert_array_equal(tail.apply_bbox_deltas(np.broadcast_to(a, d.shape), d, z["dc/mean"], z["dc/std"]),
                                  z["dc/decoded_mean_std"])
    np.testing.assert_array_equal(tail.clip_boxes(tuple(z["dc/image_shape"]), z["dc/decoded"]), z["dc/clipped"])


@pytest.mark.skipif(not os.path.exists("/root/reference/FilterDetections.py"),
                    reason="build container only: executes the reference's FilterDetections.py from where it lies")
def test_oracle_tail_against_the_reference_executed_live_on_random_cases():
    """Beyond the committed fixtures: 24 random configurations (classes, sizes, thresholds, nms / class-specific
    switches, max_detections below and above the candidate count) through the reference's own FilterDetections.py
    (numpy stand-in for its TensorFlow ops, tests/golden/tf_tail_stub.py) and through the oracle: identical."""
    import importlib.util
    import sys
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "tensorflow" or k.startswith("tensorflow.")}
    sys.path.insert(0, golden)
    try:
        import tf_tail_stub as stub
        stub.install()
        spec = importlib.util.spec_from_file_location("ref_filter_detections", "/root/reference/FilterDetections.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        rng = np.random.default_rng(20261019)
        seen_detections = seen_padding = 0
        for case in range(24):
            B, N, C = int(rng.integers(1, 4)), int(rng.integers(1, 500)), int(rng.integers(1, 8))
            span, wmax = float(rng.uniform(50, 600)), float(rng.uniform(5, 200))
            xy = rng.uniform(0, span, (B, N, 2)).astype(np.float32)
            boxes = np.concatenate([xy, xy + rng.uniform(1, wmax, (B, N, 2)).astype(np.float32)], -1)
            v = np.linspace(0, 1, B * N * C + 2, dtype=np.float64)[1:-1].astype(np.float32)
            cls = rng.permutation(v).reshape(B, N, C)
            kw = dict(nms=bool(rng.integers(0, 2)), class_specific_filter=bool(rng.integers(0, 2)),
                      nms_threshold=float(rng.choice([0.05, 0.3, 0.5, 0.8])),
                      score_threshold=float(rng.choice([0.0, 0.3, 0.7, 0.95])),
                      max_detections=int(rng.choice([1, 7, 50, 300])))
            want = ref.FilterDetections(**kw).call([stub.t(boxes), stub.t(cls)])
            got = tail.filter_detections_batch(boxes, cls, **kw)
            for g, w, nm in zip(got, want, ("boxes", "scores", "labels")):
                np.testing.assert_array_equal(g, np.asarray(w), err_msg="case %d %s %r" % (case, nm, kw))
            seen_detections += int((got[2] >= 0).sum())
            seen_padding += int((got[2] < 0).sum())
        assert seen_detections > 500 and seen_padding > 500
    finally:
        sys.path.remove(golden)
        for k in [k for k in sys.modules if k == "tensorflow" or k.startswith("tensorflow.")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})
        sys.modules.pop("tf_tail_stub", None)
