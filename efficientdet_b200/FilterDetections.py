"""FilterDetections layer + filter_detections / filter_by_score_and_nms -- mirror of the
reference's FilterDetections.py (:5-34, :37-118, :121-234) on effdet_filter_detections
(csrc/tail.cu): score threshold -> per-class sort -> greedy NMS -> cross-class top-k -> pad."""
import numpy as np
import torch

from . import _lib
from ._tensor import as_device, device, give_back
from .keras_compat import Layer

_ws_cache = {}


def _launch(boxes, classification, class_specific, score_threshold, max_detections, iou_threshold, nms, cap,
            want_indices=False):
    """One attempt with candidate capacity `cap`: enqueues the tail, no synchronisation.
    -> (out_boxes, out_scores, out_labels, out_indices | None, status (4,) i32 device tensor);
    status[0] != 0: the candidates did not fit, status[1] = capacity needed."""
    B, N, C = classification.shape
    dev = boxes.device
    out_b = torch.empty((B, max_detections, 4), dtype=torch.float32, device=dev)
    out_s = torch.empty((B, max_detections), dtype=torch.float32, device=dev)
    out_l = torch.empty((B, max_detections), dtype=torch.int32, device=dev)
    out_i = torch.empty((B, max_detections), dtype=torch.int32, device=dev) if want_indices else None
    lib = _lib.load()
    nbytes = lib.effdet_filter_detections_workspace_size(B, N, C, cap, max_detections)
    key = (dev.index, "ws")
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    _lib.call("effdet_filter_detections", boxes.data_ptr(), classification.data_ptr(), B, N, C,
              float(score_threshold), float(iou_threshold), int(max_detections),
              int(bool(class_specific)), int(bool(nms)), ws.data_ptr(), ws.numel(), cap,
              out_b.data_ptr(), out_s.data_ptr(), out_l.data_ptr(), _lib.ptr(out_i),
              status.data_ptr(), _lib.stream_ptr())
    return out_b, out_s, out_l, out_i, status


def _capacity(B, N, C, S, cand_capacity=None):
    worst = B * N * S
    cap = cand_capacity if cand_capacity is not None else _ws_cache.get(
        ("cap", B, N, C, S), min(worst, max(1 << 16, B * 8192)))
    return max(1, min(cap, worst)), worst


def _grow_capacity(B, N, C, S, status_host):
    """Remembers the capacity a failed attempt asked for; returns it."""
    worst = B * N * S
    need = int(status_host[1])
    cap = worst if need >= 0x7fffffff else min(worst, need + need // 8 + 1024)
    _ws_cache[("cap", B, N, C, S)] = cap
    return cap


def _run(boxes, classification, class_specific, score_threshold, max_detections, iou_threshold,
         nms, cand_capacity=None, want_indices=False):
    """boxes (B,N,4) f32 cuda, classification (B,N,C) f32 cuda -> (boxes, scores, labels) cuda."""
    B, N, C = classification.shape
    dev = boxes.device
    if B == 0:
        out_b = torch.empty((B, max_detections, 4), dtype=torch.float32, device=dev)
        out_s = torch.empty((B, max_detections), dtype=torch.float32, device=dev)
        out_l = torch.empty((B, max_detections), dtype=torch.int32, device=dev)
        out_i = torch.empty((B, max_detections), dtype=torch.int32, device=dev) if want_indices else None
        return (out_b, out_s, out_l, out_i) if want_indices else (out_b, out_s, out_l)
    S = C if class_specific else 1
    cap, _ = _capacity(B, N, C, S, cand_capacity)
    while True:
        out_b, out_s, out_l, out_i, status = _launch(boxes, classification, class_specific, score_threshold,
                                                     max_detections, iou_threshold, nms, cap, want_indices)
        st = status.cpu()
        if int(st[0]) == 0:
            return (out_b, out_s, out_l, out_i) if want_indices else (out_b, out_s, out_l)
        cap = _grow_capacity(B, N, C, S, st)


def filter_detections(boxes, classification, class_specific_filter=True, score_threshold=0.01,
                      max_detections=300, iou_threshold=0.5):
    """Single image: boxes (N,4), classification (N,C) -> [boxes (max,4), scores, labels i32]."""
    b, hb = as_device(boxes)
    c, hc = as_device(classification)
    ob, os_, ol = _run(b[None], c[None], class_specific_filter, score_threshold, max_detections,
                       iou_threshold, iou_threshold > 0)
    host = hb and hc
    return [give_back(ob[0], host), give_back(os_[0], host), give_back(ol[0], host)]


def filter_by_score_and_nms(scores_, labels_, score_threshold, boxes, max_detections,
                            iou_threshold):
    """Indices above the threshold (+ NMS) as (num_keeps, 2) int64 [index, label] rows --
    FilterDetections.py:5-34.  With NMS the rows are in selection (descending-score) order;
    without (iou_threshold <= 0) in ascending index order, like tf.where."""
    s, _ = as_device(scores_)
    b, _ = as_device(boxes)
    n = int(s.shape[0])
    labels = np.asarray(labels_.cpu() if isinstance(labels_, torch.Tensor) else labels_, np.int64)
    if n == 0:
        return np.zeros((0, 2), np.int64)
    do_nms = iou_threshold > 0
    k = max(1, int(max_detections)) if do_nms else n
    _, _, _, idx = _run(b.reshape(1, n, 4), s.reshape(1, n, 1), True, score_threshold, k,
                        iou_threshold, do_nms, want_indices=True)
    keep = idx[0].cpu().numpy().astype(np.int64)
    keep = keep[keep >= 0]
    if not do_nms:
        keep = np.sort(keep)
    return np.stack([keep, labels[keep]], axis=1).reshape(-1, 2)


class FilterDetections(Layer):
    """Keras-style layer: score threshold, per-class NMS and top-k over a batch."""

    def __init__(self, nms=True, class_specific_filter=True, nms_threshold=0.5,
                 score_threshold=0.01, max_detections=300, parallel_iterations=32, **kwargs):
        self.nms = nms
        self.class_specific_filter = class_specific_filter
        self.nms_threshold = nms_threshold
        self.score_threshold = score_threshold
        self.max_detections = max_detections
        self.parallel_iterations = parallel_iterations   # kept for API parity; the whole
        super(FilterDetections, self).__init__(**kwargs)  # batch runs in one launch sequence

    def call(self, inputs, **kwargs):
        boxes, hb = as_device(inputs[0])
        classification, hc = as_device(inputs[1])
        if not self.nms:
            self.nms_threshold = 0        # FilterDetections.py:170-171
        ob, os_, ol = _run(boxes, classification, self.class_specific_filter, self.score_threshold,
                           self.max_detections, self.nms_threshold,
                           bool(self.nms) and self.nms_threshold > 0)     # iou_threshold <= 0 skips NMS (:11)
        host = hb and hc
        return [give_back(ob, host), give_back(os_, host), give_back(ol, host)]

    def compute_output_shape(self, input_shape):
        return [(input_shape[0][0], self.max_detections, 4),
                (input_shape[1][0], self.max_detections),
                (input_shape[1][0], self.max_detections)]

    def compute_mask(self, inputs, mask=None):
        return (len(inputs) + 1) * [None]

    def get_config(self):
        config = super(FilterDetections, self).get_config()
        config.update({"nms": self.nms, "class_specific_filter": self.class_specific_filter,
                       "nms_threshold": self.nms_threshold,
                       "score_threshold": self.score_threshold,
                       "max_detections": self.max_detections,
                       "parallel_iterations": self.parallel_iterations})
        return config
