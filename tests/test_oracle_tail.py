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
