"""Generates tests/golden/filter_detections.npz by EXECUTING THE REFERENCE's own FilterDetections.py
(/root/reference/FilterDetections.py:5-190, imported unmodified) through tests/golden/tf_tail_stub.py's numpy stand-in
for the TensorFlow ops it calls.  The reference's control flow -- thresholding, per-class loop, labels, index gathers,
concatenation order, top-k, -1 padding, casts, the layer's `nms=False -> iou_threshold = 0`, map_fn over the batch --
is therefore the reference's; tf.image.non_max_suppression, tf.nn.top_k and tf.where are restated in the stub from
TensorFlow's documented behaviour.  Scores are distinct (TF's tie order is not pinned).
Build container only.  Run:  python tests/golden/make_golden_filter.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_tail_stub as stub  # noqa: E402

stub.install()
sys.path.insert(0, "/root/reference")
import FilterDetections as ref  # noqa: E402  (reference FilterDetections.py)

assert ref.__file__.startswith("/root/reference/")
rng = np.random.default_rng(4242)


def boxes_(n, span, wmax):
    xy = rng.uniform(0, span, (n, 2)).astype(np.float32)
    wh = rng.uniform(2, wmax, (n, 2)).astype(np.float32)
    return np.concatenate([xy, xy + wh], 1).astype(np.float32)


def scores_(shape):
    n = int(np.prod(shape))
    v = np.linspace(0, 1, n + 2, dtype=np.float64)[1:-1].astype(np.float32)
    assert len(np.unique(v)) == n
    return rng.permutation(v).reshape(shape)


cases = {
    # name: (B, N, C, span, wmax, layer kwargs)
    "pad": (2, 300, 4, 400, 60, dict(score_threshold=0.97, max_detections=40)),                 # fewer than max -> -1 padding
    "topk": (2, 500, 5, 300, 120, dict(score_threshold=0.6, max_detections=50)),                # far more than max -> top-k
    "class_cap": (1, 800, 2, 4000, 10, dict(score_threshold=0.5, max_detections=30)),           # per-class NMS cap reached
    "max_class": (2, 400, 6, 300, 100, dict(score_threshold=0.7, class_specific_filter=False, max_detections=60)),
    "no_nms": (2, 300, 3, 200, 150, dict(score_threshold=0.8, nms=False, max_detections=100)),
    "no_nms_max_class": (1, 300, 3, 200, 150, dict(score_threshold=0.8, nms=False, class_specific_filter=False,
                                                   max_detections=25)),
    "tight_iou": (1, 400, 3, 200, 120, dict(score_threshold=0.5, nms_threshold=0.1, max_detections=300)),
    "nothing": (2, 100, 3, 200, 50, dict(score_threshold=2.0, max_detections=20)),
}
out = {}
for name, (B, N, C, span, wmax, kw) in cases.items():
    boxes = np.stack([boxes_(N, span, wmax) for _ in range(B)])
    if name == "tight_iou":                      # degenerate boxes: zero area and inverted corners
        boxes[:, ::7, 2] = boxes[:, ::7, 0]
        boxes[:, ::11, [0, 2]] = boxes[:, ::11, [2, 0]]
    cls = scores_((B, N, C))
    layer = ref.FilterDetections(**kw)
    got = layer.call([stub.t(boxes), stub.t(cls)])
    assert layer.compute_output_shape([boxes.shape, cls.shape])[0] == (B, layer.max_detections, 4)
    out[name + "/boxes_in"], out[name + "/cls_in"] = boxes, cls
    out[name + "/kw"] = np.array(repr(sorted(kw.items())))
    out[name + "/boxes"] = np.asarray(got[0], np.float32)
    out[name + "/scores"] = np.asarray(got[1], np.float32)
    out[name + "/labels"] = np.asarray(got[2], np.int32)
    print(name, out[name + "/boxes"].shape, int((out[name + "/labels"] >= 0).sum()), "detections")
# the function form on one image (FilterDetections.py:37-118) and filter_by_score_and_nms alone (:5-34)
b1, c1 = boxes_(200, 150, 80), scores_((200, 4))
r = ref.filter_detections(stub.t(b1), stub.t(c1), score_threshold=0.6, max_detections=35, iou_threshold=0.4)
out["fn/boxes_in"], out["fn/cls_in"] = b1, c1
out["fn/boxes"], out["fn/scores"], out["fn/labels"] = (np.asarray(r[0], np.float32), np.asarray(r[1], np.float32),
                                                       np.asarray(r[2], np.int32))
lab = rng.integers(0, 4, 200).astype(np.int64)
idx = ref.filter_by_score_and_nms(stub.t(c1[:, 1]), stub.t(lab), 0.5, stub.t(b1), 20, 0.45)
out["fsn/labels_in"], out["fsn/indices"] = lab, np.asarray(idx, np.int64)
# decode + clip: the reference's own apply_bbox_deltas (RegressBoxes.py:126-164) and ClipBoxes.call (ClipBoxes.py:9-24)
import contextlib  # noqa: E402
import io  # noqa: E402

import ClipBoxes as ref_clip  # noqa: E402
import RegressBoxes as ref_reg  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.anchors import anchors_for_shape  # noqa: E402  (inputs only: any boxes would do)

anchors = anchors_for_shape((256, 384)).astype(np.float32)[None]
deltas = rng.normal(0, 0.6, (3, anchors.shape[1], 4)).astype(np.float32)
dec = ref_reg.apply_bbox_deltas(stub.t(np.broadcast_to(anchors, deltas.shape).copy()), stub.t(deltas))
mean, std = np.array([0.1, -0.1, 0.05, 0.0], np.float32), np.array([0.1, 0.2, 0.3, 0.25], np.float32)
dec2 = ref_reg.apply_bbox_deltas(stub.t(np.broadcast_to(anchors, deltas.shape).copy()), stub.t(deltas), mean, std)
with contextlib.redirect_stdout(io.StringIO()):          # ClipBoxes.call prints its inputs (ClipBoxes.py:12-14)
    clipped = ref_clip.ClipBoxes().call([stub.t(np.zeros((3, 256, 384, 3), np.float32)), dec])
out.update({"dc/anchors": anchors, "dc/deltas": deltas, "dc/decoded": np.asarray(dec, np.float32),
            "dc/mean": mean, "dc/std": std, "dc/decoded_mean_std": np.asarray(dec2, np.float32),
            "dc/clipped": np.asarray(clipped, np.float32), "dc/image_shape": np.array([3, 256, 384, 3])})
assert out["dc/decoded"].dtype == np.float32 and (out["dc/clipped"] != out["dc/decoded"]).any()
np.savez_compressed(os.path.join(HERE, "filter_detections.npz"), **out)
print("wrote", len(out), "arrays")
