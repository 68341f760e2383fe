"""Mapping detections back from the letterboxed network input to the original image
(inference.py:72-85, eval/common.py:91-97, predict.py:107-121)."""
import numpy as np


def unletterbox_boxes(boxes, scale, offset_h, offset_w, height, width):
    """boxes (..., 4) as (x1, y1, x2, y2) in network-input pixels -> original-image pixels:
    subtract the letterbox offsets, divide by the resize scale, clip to [0, w-1] x [0, h-1]."""
    boxes = np.array(boxes, dtype=np.float32, copy=True)
    boxes[..., [0, 2]] -= offset_w
    boxes[..., [1, 3]] -= offset_h
    boxes /= scale
    boxes[..., 0] = np.clip(boxes[..., 0], 0, width - 1)
    boxes[..., 2] = np.clip(boxes[..., 2], 0, width - 1)
    boxes[..., 1] = np.clip(boxes[..., 1], 0, height - 1)
    boxes[..., 3] = np.clip(boxes[..., 3], 0, height - 1)
    return boxes


def select_detections(boxes, scores, labels, score_threshold):
    """inference.py:87-95: keep the detections of one image whose score exceeds the threshold."""
    idx = np.where(scores > score_threshold)[0]
    return boxes[idx], scores[idx], labels[idx]
