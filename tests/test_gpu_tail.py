"""GPU parity (bit-exact) of the detection tail and geometry kernels against the oracle,
through the reference-shaped Python API -> C ABI -> CUDA."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _boxes(rng, n, span=512, wmax=160):
    xy = rng.uniform(0, span, (n, 2)).astype(np.float32)
    wh = rng.uniform(4, wmax, (n, 2)).astype(np.float32)
    return np.concatenate([xy, xy + wh], 1).astype(np.float32)


def _distinct_scores(rng, shape, lo=0.0, hi=1.0):
    n = int(np.prod(shape))
    v = np.linspace(lo, hi, n + 2, dtype=np.float64)[1:-1].astype(np.float32)
    assert len(np.unique(v)) == n
    return rng.permutation(v).reshape(shape)


# ---------------------------------------------------------------- reference test vectors
def test_apply_bbox_deltas_reference_vector():
    from efficientdet_b200 import RegressBoxes as RB
    boxes = np.array([[[0, 0, 1, 1], [0.5, 0.5, 0.6, 0.6]]], "float32")
    deltas = np.array([[[0.1, 0.1, 0.1, 0.1], [-0.2, -0.2, 0.2, 0.2]]], "float32")
    want = np.array([[[0.02, 0.02, 1.02, 1.02], [0.496, 0.496, 0.604, 0.604]]])
    np.testing.assert_array_almost_equal(RB.apply_bbox_deltas(boxes, deltas), want)
    layer = RB.RegressBoxes(anchor_shape=boxes.shape)
    layer.set_anchors(boxes)
    np.testing.assert_array_almost_equal(layer([deltas]), want)
    np.testing.assert_array_almost_equal(RB.RegressBoxes()([boxes, deltas]), want)
    mean, std = np.array([0.5] * 4), np.array([0.1] * 4)
    lay = RB.RegressBoxes(mean=mean, std=std)
    np.testing.assert_array_equal(lay.mean, mean)
    np.testing.assert_array_equal(lay.std, std)
    assert RB.RegressBoxes().compute_output_shape([[10], [10]]) == [10]
    with pytest.raises(ValueError):
        RB.RegressBoxes(mean=3.0)


def test_clip_boxes_reference_vector():
    from efficientdet_b200 import ClipBoxes as CB
    boxes = np.array([[[-0.1, 0, 1.3, 1], [10, 0, 210, 300], [-100, -0.5, 0.6, 180]]], np.float32)
    want = np.array([[[0, 0, 1.3, 1], [10, 0, 199, 199], [0, 0, 0.6, 180]]], np.float32)
    img = np.ones((32, 200, 200, 3))
    np.testing.assert_array_equal(CB.ClipBoxes().call([img, boxes]), want)
    assert CB.ClipBoxes().compute_output_shape([[10], [10]]) == [10]


def test_filter_by_score_and_nms_reference_vector():
    from efficientdet_b200 import FilterDetections as FD
    boxes = np.array([[0, 0, 1, 1], [0.5, 0.5, 0.6, 0.6], [0.1, 0.1, 0.6, 0.6]], "float32")
    scores = np.array([.6, .2, .1], "float32")
    labels = np.array([1, 2, 1], "int64")
    got = FD.filter_by_score_and_nms(scores, labels, .12, boxes, 3, .5)
    np.testing.assert_array_almost_equal(got, np.array([[0, 1], [1, 2]]))
    got = FD.filter_by_score_and_nms(scores, labels, .12, boxes, 3, 0)
    np.testing.assert_array_equal(got, np.array([[0, 1], [1, 2]]))


# ---------------------------------------------------------------- decode / clip parity
@pytest.mark.parametrize("B,N,shared", [(1, 49104, True), (3, 1000, False), (2, 7, True), (4, 0, True)])
def test_regress_clip_bit_exact(B, N, shared):
    from efficientdet_b200 import RegressBoxes as RB, ClipBoxes as CB
    from oracle import tail
    rng = np.random.default_rng(B * 1000 + N)
    anchors = _boxes(rng, (1 if shared else B) * N).reshape((1 if shared else B), N, 4)
    deltas = rng.normal(0, 0.5, (B, N, 4)).astype(np.float32)
    got = RB.apply_bbox_deltas(anchors, deltas)
    want = tail.apply_bbox_deltas(np.broadcast_to(anchors, (B, N, 4)), deltas)
    assert np.array_equal(got, want)
    got_c = CB.clip_boxes((B, 384, 512, 3), got)
    assert np.array_equal(got_c, tail.clip_boxes((B, 384, 512, 3), want))


def test_fused_regress_clip_matches_unfused():
    import torch
    from efficientdet_b200 import _lib
    from oracle import tail
    rng = np.random.default_rng(3)
    B, N = 2, 5000
    anchors = _boxes(rng, N).reshape(1, N, 4)
    deltas = rng.normal(0, 0.7, (B, N, 4)).astype(np.float32)
    a = torch.from_numpy(anchors).cuda(); d = torch.from_numpy(deltas).cuda()
    out = torch.empty_like(d)
    f4 = _lib.c_float * 4
    _lib.call("effdet_regress_clip_boxes", a.data_ptr(), 0, d.data_ptr(), f4(0, 0, 0, 0),
              f4(.2, .2, .2, .2), B, N, 512.0, 512.0, out.data_ptr(), _lib.stream_ptr())
    want = tail.clip_boxes((B, 512, 512, 3), tail.apply_bbox_deltas(np.broadcast_to(anchors, (B, N, 4)), deltas))
    assert np.array_equal(out.cpu().numpy(), want)


# ---------------------------------------------------------------- FilterDetections parity
def _check_fd(boxes, cls, **kw):
    from efficientdet_b200.FilterDetections import FilterDetections
    from oracle import tail
    layer = FilterDetections(**kw)
    gb, gs, gl = layer([boxes, cls])
    wb, ws, wl = tail.filter_detections_batch(
        boxes, cls, nms=kw.get("nms", True), class_specific_filter=kw.get("class_specific_filter", True),
        nms_threshold=kw.get("nms_threshold", 0.5), score_threshold=kw.get("score_threshold", 0.01),
        max_detections=kw.get("max_detections", 300))
    assert gl.dtype == np.int32 and gb.dtype == np.float32
    assert np.array_equal(gs, ws), (np.argwhere(gs != ws)[:5])
    assert np.array_equal(gl, wl)
    assert np.array_equal(gb, wb)
    return gs


@pytest.mark.parametrize("B,N,C,thr", [(2, 3000, 5, 0.9), (1, 20000, 20, 0.99), (3, 1200, 90, 0.97)])
def test_filter_detections_bit_exact(B, N, C, thr):
    rng = np.random.default_rng(N + C)
    boxes = np.stack([_boxes(rng, N, 300, 120) for _ in range(B)])
    cls = _distinct_scores(rng, (B, N, C))
    s = _check_fd(boxes, cls, score_threshold=thr)
    assert (s >= 0).sum() > 0


def test_filter_detections_large_segments_and_cap():
    """> 2048 candidates per class (global-memory sort path) and per-class 300 cap."""
    rng = np.random.default_rng(77)
    B, N, C = 1, 12000, 3
    boxes = np.stack([_boxes(rng, N, 2000, 40)])
    cls = _distinct_scores(rng, (B, N, C))
    _check_fd(boxes, cls, score_threshold=0.5, max_detections=300)
    _check_fd(boxes, cls, score_threshold=0.5, max_detections=50)


def test_filter_detections_max_detections_2048():
    """The largest max_detections the NMS path accepts: the selected-box list (32 KiB) + the small-segment kernel's
    static shared memory pass 48 KiB (the launch opts in)."""
    rng = np.random.default_rng(2048)
    B, N, C = 1, 2900, 2                  # ~2030 candidates per class: tiny boxes far apart, nearly all survive NMS
    boxes = np.stack([_boxes(rng, N, 4000, 12)])
    cls = _distinct_scores(rng, (B, N, C))
    s = _check_fd(boxes, cls, score_threshold=0.3, max_detections=2048)
    assert (s >= 0).sum() == 2048
    _check_fd(boxes, cls, score_threshold=0.3, max_detections=1900)


def test_filter_detections_variants():
    rng = np.random.default_rng(5)
    B, N, C = 2, 2500, 7
    boxes = np.stack([_boxes(rng, N, 200, 100) for _ in range(B)])
    cls = _distinct_scores(rng, (B, N, C))
    _check_fd(boxes, cls, score_threshold=0.8, nms=False)
    _check_fd(boxes, cls, score_threshold=0.8, class_specific_filter=False)
    _check_fd(boxes, cls, score_threshold=0.8, class_specific_filter=False, nms=False, max_detections=100)
    _check_fd(boxes, cls, score_threshold=0.8, nms_threshold=0.3, max_detections=40)
    # nothing above the threshold -> all padding
    s = _check_fd(boxes, cls, score_threshold=2.0)
    assert np.all(s == -1)
    # degenerate (zero-area / inverted) boxes follow TF's IoU rules
    boxes2 = boxes.copy()
    boxes2[:, ::7, 2] = boxes2[:, ::7, 0]
    boxes2[:, ::11, [0, 2]] = boxes2[:, ::11, [2, 0]]
    _check_fd(boxes2, cls, score_threshold=0.8)


def test_filter_detections_nms_stress_5000():
    """BASELINE config 5 shape: ~5000 pre-NMS (anchor,class) pairs per image, C = 90."""
    from oracle import anchors as oa, tail
    rng = np.random.default_rng(99)
    S, B, C = 512, 2, 90
    anchors = oa.anchors_for_shape((S, S)).astype(np.float32)
    N = anchors.shape[0]
    deltas = rng.normal(0, 0.5, (B, N, 4)).astype(np.float32)
    boxes = tail.clip_boxes((B, S, S, 3), tail.apply_bbox_deltas(anchors[None], deltas))
    cls = rng.uniform(0, 0.0099, (B, N, C)).astype(np.float32)
    hot = _distinct_scores(rng, (B, 5000), 0.05, 1.0)
    for b in range(B):
        slots = rng.choice(N * C, 5000, replace=False)
        cls[b].reshape(-1)[slots] = hot[b]
    _check_fd(boxes, cls, score_threshold=0.01)


# ---------------------------------------------------------------- geometry parity
def test_compute_overlap_bit_exact(golden):
    from efficientdet_b200.utils.compute_overlap import compute_overlap
    assert np.array_equal(compute_overlap(golden["ov_boxes"], golden["ov_query"]), golden["ov_result"])
    kat = compute_overlap(np.array([[0, 0, 10, 10], [5, 5, 15, 15]], np.float64),
                          np.array([[0, 0, 10, 10]], np.float64))
    assert np.array_equal(kat, golden["ov_kat"])
    with pytest.raises(ValueError):
        compute_overlap(np.zeros((2, 4), np.float32), np.zeros((1, 4)))


def test_anchor_targets_bit_exact(golden):
    from efficientdet_b200.utils import anchors as A
    a = A.anchors_for_shape((128, 128))
    shapes = [tuple(s) for s in golden["tg_img_shapes"]]
    ann = [{"bboxes": golden["tg_bboxes_%d" % i], "labels": golden["tg_labels_%d" % i]}
           for i in range(len(shapes))]
    reg, lab = A.anchor_targets_bbox(a, [np.zeros(s) for s in shapes], ann, 6)
    assert np.array_equal(reg, golden["tg_regression"])
    assert np.array_equal(lab, golden["tg_labels"])
    pos, ign, arg = A.compute_gt_annotations(a, golden["tg_bboxes_0"])
    from oracle import anchors as oa
    p2, i2, a2 = oa.compute_gt_annotations(a, golden["tg_bboxes_0"])
    assert np.array_equal(pos, p2) and np.array_equal(ign, i2) and np.array_equal(arg, a2)


def test_anchor_targets_kat_512(golden):
    import hashlib
    from efficientdet_b200.utils import anchors as A
    a = A.anchors_for_shape((512, 512))
    reg, lab = A.anchor_targets_bbox(
        a, [np.zeros((512, 512, 3))],
        [{"bboxes": np.array([[100, 120, 300, 360], [10, 10, 60, 80]], np.float32),
          "labels": np.array([3, 7], np.float32)}], 20)
    assert (reg[0, :, 4] == 1).sum() == 68 and (reg[0, :, 4] == -1).sum() == 164
    assert hashlib.sha256(reg.tobytes()).hexdigest() == str(golden["kat_reg_sha"])
    assert hashlib.sha256(lab.tobytes()).hexdigest() == str(golden["kat_lab_sha"])
    with pytest.raises(AssertionError):
        A.anchor_targets_bbox(a, [], [], 20)


# ---------------------------------------------------------------- fixtures from the EXECUTED reference
@pytest.mark.parametrize("case", ["pad", "topk", "class_cap", "max_class", "no_nms", "no_nms_max_class", "tight_iou",
                                  "nothing"])
def test_filter_detections_matches_executed_reference(case):
    """tests/golden/filter_detections.npz holds the outputs of the reference's own FilterDetections.py, executed
    unmodified over a numpy stand-in for its TensorFlow ops (tests/golden/make_golden_filter.py): the CUDA tail must
    reproduce them bit for bit (boxes, scores, labels, padding)."""
    import ast
    import os
    from efficientdet_b200.FilterDetections import FilterDetections
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filter_detections.npz"))
    kw = dict(ast.literal_eval(str(z[case + "/kw"])))
    gb, gs, gl = FilterDetections(**kw)([z[case + "/boxes_in"], z[case + "/cls_in"]])
    assert gl.dtype == np.int32
    assert np.array_equal(gl, z[case + "/labels"])
    assert np.array_equal(gs, z[case + "/scores"])
    assert np.array_equal(gb, z[case + "/boxes"])


def test_decode_and_clip_match_executed_reference():
    """The reference's own apply_bbox_deltas (RegressBoxes.py:126-164) and ClipBoxes.call (ClipBoxes.py:9-24),
    executed by tests/golden/make_golden_filter.py on 3 x 18 414 anchors: the CUDA decode / clip reproduce the float32
    results bit for bit."""
    import os
    from efficientdet_b200 import RegressBoxes as RB, ClipBoxes as CB
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filter_detections.npz"))
    got = RB.apply_bbox_deltas(z["dc/anchors"], z["dc/deltas"])
    assert np.array_equal(got, z["dc/decoded"])
    assert np.array_equal(CB.clip_boxes(tuple(int(v) for v in z["dc/image_shape"]), got), z["dc/clipped"])
