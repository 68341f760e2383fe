"""Oracle: detection tail in numpy float32 (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates
  /root/reference/RegressBoxes.py:126-164   apply_bbox_deltas
  /root/reference/ClipBoxes.py:9-24         ClipBoxes.call
  /root/reference/FilterDetections.py:5-34  filter_by_score_and_nms
  /root/reference/FilterDetections.py:37-118 filter_detections
  /root/reference/FilterDetections.py:155-190 FilterDetections.call (map over batch)
TensorFlow ops whose arithmetic is not in the reference tree (pinned
tensorflow 1.15.0 / 2.0.0, requirements.txt:21, Dockerfile:1) are restated from
their published behaviour: tf.where (ascending row-major), non_max_suppression
V3 (greedy, strict >, float32 IoU without +1, corners min/max-normalised,
zero-area => IoU 0), tf.nn.top_k (descending, ties -> lower index).
Equal-score order inside NMS is unspecified in the pinned TF versions; this
oracle takes the lower index first and the test harness forbids score ties.
"""
import numpy as np

F = np.float32


def apply_bbox_deltas(boxes, deltas, mean=(0, 0, 0, 0), std=(0.2, 0.2, 0.2, 0.2)):
    boxes = np.asarray(boxes, F)
    deltas = np.asarray(deltas, F)
    mean = np.asarray(mean, F)
    std = np.asarray(std, F)
    w = boxes[:, :, 2] - boxes[:, :, 0]
    h = boxes[:, :, 3] - boxes[:, :, 1]
    scales = np.stack([w, h, w, h], axis=2)
    nd = deltas * std + mean            # two separately rounded float32 ops
    return (boxes + nd * scales).astype(F)


def clip_boxes(image_shape, boxes):
    """image_shape: (B,H,W,C) tuple; x clipped to [0,W-1], y to [0,H-1]."""
    boxes = np.asarray(boxes, F)
    H, W = F(image_shape[1]), F(image_shape[2])
    out = np.empty_like(boxes)
    out[:, :, 0] = np.clip(boxes[:, :, 0], F(0), W - F(1))
    out[:, :, 1] = np.clip(boxes[:, :, 1], F(0), H - F(1))
    out[:, :, 2] = np.clip(boxes[:, :, 2], F(0), W - F(1))
    out[:, :, 3] = np.clip(boxes[:, :, 3], F(0), H - F(1))
    return out


def _iou(b, i, rows):
    """float32 IoU of box i against boxes[rows] the way TF's NMS kernel does it."""
    bi = b[i]
    y0i, y1i = min(bi[0], bi[2]), max(bi[0], bi[2])
    x0i, x1i = min(bi[1], bi[3]), max(bi[1], bi[3])
    r = b[rows]
    y0 = np.minimum(r[:, 0], r[:, 2]); y1 = np.maximum(r[:, 0], r[:, 2])
    x0 = np.minimum(r[:, 1], r[:, 3]); x1 = np.maximum(r[:, 1], r[:, 3])
    ai = F(y1i - y0i) * F(x1i - x0i)
    aj = (y1 - y0) * (x1 - x0)
    ih = np.maximum(np.minimum(y1i, y1) - np.maximum(y0i, y0), F(0))
    iw = np.maximum(np.minimum(x1i, x1) - np.maximum(x0i, x0), F(0))
    inter = (ih * iw).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / ((ai + aj).astype(F) - inter)
    iou = np.where((ai <= 0) | (aj <= 0), F(0), iou)
    return iou.astype(F)


def non_max_suppression(boxes, scores, max_output_size, iou_threshold):
    boxes = np.asarray(boxes, F).reshape(-1, 4)
    scores = np.asarray(scores, F)
    order = np.lexsort((np.arange(scores.shape[0]), -scores.astype(np.float64)))
    thr = F(iou_threshold)
    selected = []
    for i in order:
        if len(selected) >= max_output_size:
            break
        if selected:
            if np.any(_iou(boxes, i, np.asarray(selected)) > thr):
                continue
        selected.append(int(i))
    return np.asarray(selected, np.int64)


def filter_by_score_and_nms(scores, labels, score_threshold, boxes, max_detections, iou_threshold):
    scores = np.asarray(scores, F)
    idx = np.nonzero(scores > F(score_threshold))[0].astype(np.int64)
    if iou_threshold > 0:
        keep = non_max_suppression(np.asarray(boxes, F)[idx], scores[idx], max_detections,
                                   iou_threshold)
        idx = idx[keep]
    lab = np.asarray(labels, np.int64)[idx]
    return np.stack([idx, lab], axis=1).reshape(-1, 2)


def filter_detections(boxes, classification, class_specific_filter=True, score_threshold=0.01,
                      max_detections=300, iou_threshold=0.5):
    boxes = np.asarray(boxes, F)
    classification = np.asarray(classification, F)
    n, C = classification.shape
    if class_specific_filter:
        parts = []
        for c in range(C):
            parts.append(filter_by_score_and_nms(classification[:, c], np.full(n, c, np.int64),
                                                 score_threshold, boxes, max_detections,
                                                 iou_threshold))
        indices = np.concatenate(parts, axis=0) if parts else np.zeros((0, 2), np.int64)
    else:
        scores = classification.max(axis=1) if C else np.zeros((n,), F)
        labels = classification.argmax(axis=1) if C else np.zeros((n,), np.int64)
        indices = filter_by_score_and_nms(scores, labels, score_threshold, boxes, max_detections,
                                          iou_threshold)
    scores = classification[indices[:, 0], indices[:, 1]]
    labels = indices[:, 1]
    k = min(max_detections, scores.shape[0])
    top = np.lexsort((np.arange(scores.shape[0]), -scores.astype(np.float64)))[:k]
    sel_scores = scores[top]
    sel_boxes = boxes[indices[top, 0]]
    sel_labels = labels[top]
    pad = max(0, max_detections - k)
    out_b = np.concatenate([sel_boxes, np.full((pad, 4), -1, F)], 0).astype(F)
    out_s = np.concatenate([sel_scores, np.full((pad,), -1, F)], 0).astype(F)
    out_l = np.concatenate([sel_labels, np.full((pad,), -1, np.int64)], 0).astype(np.int32)
    return out_b, out_s, out_l


def filter_detections_batch(boxes, classification, nms=True, class_specific_filter=True,
                            nms_threshold=0.5, score_threshold=0.01, max_detections=300):
    iou = nms_threshold if nms else 0
    outs = [filter_detections(boxes[i], classification[i], class_specific_filter, score_threshold,
                              max_detections, iou) for i in range(boxes.shape[0])]
    return (np.stack([o[0] for o in outs]), np.stack([o[1] for o in outs]),
            np.stack([o[2] for o in outs]))
