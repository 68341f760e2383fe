"""ctypes wrapper over oracle/overlap.c (TEST INFRASTRUCTURE -- see oracle/__init__.py)."""
import ctypes

import numpy as np

from . import build_ref

_lib = None


def _get():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_ref.build_own())
        _lib.oracle_compute_overlap.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                                ctypes.c_size_t, ctypes.c_void_p]
        _lib.oracle_anchor_targets_image.argtypes = [
            ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
            ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
            ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def compute_overlap(boxes, query):
    boxes = np.ascontiguousarray(boxes, np.float64)
    query = np.ascontiguousarray(query, np.float64)
    out = np.zeros((boxes.shape[0], query.shape[0]), np.float64)
    _get().oracle_compute_overlap(boxes.ctypes.data, boxes.shape[0], query.ctypes.data,
                                  query.shape[0], out.ctypes.data)
    return out


def anchor_targets_bbox(anchors, image_shapes, annotations_group, num_classes,
                        negative_overlap=0.4, positive_overlap=0.5):
    anchors = np.ascontiguousarray(anchors, np.float64)
    B, N = len(image_shapes), anchors.shape[0]
    reg = np.zeros((B, N, 5), np.float32)
    lab = np.zeros((B, N, num_classes + 1), np.float32)
    for i, (shp, ann) in enumerate(zip(image_shapes, annotations_group)):
        gt = np.ascontiguousarray(ann["bboxes"], np.float64).reshape(-1, 4)
        gl = np.ascontiguousarray(np.asarray(ann["labels"]).astype(np.int32))
        h, w = (float(shp[0]), float(shp[1])) if len(shp) else (-1.0, -1.0)
        _get().oracle_anchor_targets_image(anchors.ctypes.data, N, gt.ctypes.data, gl.ctypes.data,
                                           gt.shape[0], num_classes, negative_overlap,
                                           positive_overlap, h, w, reg[i].ctypes.data,
                                           lab[i].ctypes.data)
    return reg, lab
