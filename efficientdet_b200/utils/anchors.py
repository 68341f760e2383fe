"""Anchor generation and anchor-target assignment -- mirror of the reference's
utils/anchors.py public surface (AnchorParameters, generate_anchors, shift, guess_shapes,
anchors_for_shape, compute_gt_annotations, anchor_targets_bbox, bbox_transform).

The anchor table is built by the library's host routine (float64, bit-exact w.r.t. the
reference); IoU / argmax / target encoding run on the GPU (csrc/targets.cu) in float64.
"""
import numpy as np
import torch

from .. import _lib
from .._tensor import as_device, device
from .compute_overlap import compute_overlap


class AnchorParameters:
    """sizes / strides per pyramid level, ratios / scales per location."""

    def __init__(self, sizes, strides, ratios, scales):
        self.sizes = sizes
        self.strides = strides
        self.ratios = ratios
        self.scales = scales

    def num_anchors(self):
        return len(self.ratios) * len(self.scales)


# keras.backend.floatx() is 'float32' in the reference (utils/anchors.py:46-52)
AnchorParameters.default = AnchorParameters(
    sizes=[32, 64, 128, 256, 512],
    strides=[8, 16, 32, 64, 128],
    ratios=np.array([0.5, 1, 2], "float32"),
    scales=np.array([2 ** 0, 2 ** (1.0 / 3.0), 2 ** (2.0 / 3.0)], "float32"),
)


def guess_shapes(image_shape, pyramid_levels):
    image_shape = np.array(image_shape[:2])
    return [(image_shape + 2 ** x - 1) // (2 ** x) for x in pyramid_levels]


def _table(level_hw, sizes, strides, ratios, scales):
    level_hw = np.ascontiguousarray(level_hw, np.int32).reshape(-1, 2)
    sizes = np.ascontiguousarray(sizes, np.int32)
    strides = np.ascontiguousarray(strides, np.int32)
    r = np.ascontiguousarray(np.asarray(ratios), np.float64)
    s = np.ascontiguousarray(np.asarray(scales), np.float64)
    rows = int((level_hw[:, 0].astype(np.int64) * level_hw[:, 1]).sum()) * len(r) * len(s)
    out = np.zeros((rows, 4), np.float64)
    _lib.call("effdet_anchors_for_shape_host", level_hw.ctypes.data, sizes.ctypes.data,
              strides.ctypes.data, level_hw.shape[0], r.ctypes.data, len(r), s.ctypes.data, len(s),
              out.ctypes.data, rows)
    return out


def generate_anchors(base_size=16, ratios=None, scales=None):
    """(len(ratios)*len(scales), 4) reference windows centred on the origin."""
    ratios = AnchorParameters.default.ratios if ratios is None else ratios
    scales = AnchorParameters.default.scales if scales is None else scales
    if int(base_size) != base_size:
        raise ValueError("base_size must be integral")
    # one 1x1 level with stride 0 puts the single cell at the origin
    return _table([[1, 1]], [int(base_size)], [0], ratios, scales)


def shift(shape, stride, anchors):
    """Replicates `anchors` (A,4) over a shape[0] x shape[1] grid of cell centres."""
    anchors = np.asarray(anchors, np.float64)
    sx = (np.arange(0, shape[1]) + 0.5) * stride
    sy = (np.arange(0, shape[0]) + 0.5) * stride
    out = np.empty((shape[0], shape[1], anchors.shape[0], 4))
    out[..., 0::2] = anchors[None, None, :, 0::2] + sx[None, :, None, None]
    out[..., 1::2] = anchors[None, None, :, 1::2] + sy[:, None, None, None]
    return out.reshape(-1, 4)


def anchors_for_shape(image_shape, pyramid_levels=None, anchor_params=None, shapes_callback=None):
    """(N,4) float64 anchors (x1,y1,x2,y2), level-major, cell row-major, (ratio, scale) minor."""
    if pyramid_levels is None:
        pyramid_levels = [3, 4, 5, 6, 7]
    if anchor_params is None:
        anchor_params = AnchorParameters.default
    if shapes_callback is None:
        shapes_callback = guess_shapes
    image_shapes = shapes_callback(image_shape, pyramid_levels)
    n = len(pyramid_levels)
    if n == 0:
        return np.zeros((0, 4))
    return _table(np.array([list(s[:2]) for s in image_shapes]), anchor_params.sizes[:n],
                  anchor_params.strides[:n], anchor_params.ratios, anchor_params.scales)


def bbox_transform(anchors, gt_boxes, mean=None, std=None):
    """Regression targets (gt - anchor) / anchor_wh, normalised by mean / std (float64)."""
    if mean is None:
        mean = np.array([0, 0, 0, 0])
    if std is None:
        std = np.array([0.2, 0.2, 0.2, 0.2])
    if isinstance(mean, (list, tuple)):
        mean = np.array(mean)
    elif not isinstance(mean, np.ndarray):
        raise ValueError("Expected mean to be a np.ndarray, list or tuple. Received: {}".format(
            type(mean)))
    if isinstance(std, (list, tuple)):
        std = np.array(std)
    elif not isinstance(std, np.ndarray):
        raise ValueError("Expected std to be a np.ndarray, list or tuple. Received: {}".format(
            type(std)))
    wh = np.stack([anchors[:, 2] - anchors[:, 0], anchors[:, 3] - anchors[:, 1]] * 2, axis=1)
    return ((gt_boxes - anchors) / wh - mean) / std


def compute_gt_annotations(anchors, annotations, negative_overlap=0.4, positive_overlap=0.5):
    """-> (positive mask, ignore mask, argmax GT index) per anchor."""
    overlaps = compute_overlap(anchors.astype(np.float64), annotations.astype(np.float64))
    argmax = np.argmax(overlaps, axis=1)
    mx = overlaps[np.arange(overlaps.shape[0]), argmax]
    positive = mx >= positive_overlap
    ignore = (mx > negative_overlap) & ~positive
    return positive, ignore, argmax


def _pack_annotations(annotations_group):
    B = len(annotations_group)
    kmax = max([int(np.asarray(a["bboxes"]).reshape(-1, 4).shape[0]) for a in annotations_group] + [1])
    gt = np.zeros((B, kmax, 4), np.float64)
    gl = np.zeros((B, kmax), np.int32)
    cnt = np.zeros((B,), np.int32)
    for i, a in enumerate(annotations_group):
        bb = np.asarray(a["bboxes"]).reshape(-1, 4)
        k = bb.shape[0]
        cnt[i] = k
        if k:
            gt[i, :k] = bb.astype(np.float64)
            gl[i, :k] = np.asarray(a["labels"]).astype(int)
    return gt, gl, cnt, kmax


def anchor_targets_device(anchors, image_shapes, annotations_group, num_classes,
                          negative_overlap=0.4, positive_overlap=0.5, dense_labels=True,
                          compact=False):
    """Device-resident variant used by the training step: returns torch CUDA tensors
    (regression (B,N,5), labels (B,N,C+1) or None, state (B,N) i8 or None, cls (B,N) i32 or None)."""
    a, _ = as_device(anchors, torch.float64)
    gt, gl, cnt, kmax = _pack_annotations(annotations_group)
    B, N = len(annotations_group), a.shape[0]
    hw = np.array([[float(s[0]), float(s[1])] if len(s) else [-1.0, -1.0] for s in image_shapes],
                  np.float64).reshape(B, 2)
    dev = device()
    gt_d = torch.from_numpy(gt).to(dev)
    gl_d = torch.from_numpy(gl).to(dev)
    cnt_d = torch.from_numpy(cnt).to(dev)
    hw_d = torch.from_numpy(hw).to(dev)
    reg = torch.empty((B, N, 5), dtype=torch.float32, device=dev)
    lab = torch.empty((B, N, num_classes + 1), dtype=torch.float32, device=dev) if dense_labels else None
    st = torch.empty((B, N), dtype=torch.int8, device=dev) if compact else None
    cl = torch.empty((B, N), dtype=torch.int32, device=dev) if compact else None
    _lib.call("effdet_anchor_targets", a.data_ptr(), N, gt_d.data_ptr(), gl_d.data_ptr(),
              cnt_d.data_ptr(), B, kmax, hw_d.data_ptr(), int(num_classes), float(negative_overlap),
              float(positive_overlap), reg.data_ptr(), _lib.ptr(lab), _lib.ptr(st), _lib.ptr(cl),
              _lib.stream_ptr())
    return reg, lab, st, cl


def anchor_targets_bbox(anchors, image_group, annotations_group, num_classes,
                        negative_overlap=0.4, positive_overlap=0.5):
    """-> (regression_batch (B,N,5) f32, labels_batch (B,N,C+1) f32) numpy, last column =
    anchor state (-1 ignore, 0 background, 1 foreground) -- utils/anchors.py:130-207."""
    assert (len(image_group) == len(annotations_group)), \
        "The length of the images and annotations need to be equal."
    assert (len(annotations_group) > 0), "No data received to compute anchor targets for."
    for annotations in annotations_group:
        assert ('bboxes' in annotations), "Annotations should contain bboxes."
        assert ('labels' in annotations), "Annotations should contain labels."
    shapes = [tuple(getattr(im, "shape", im)) for im in image_group]
    reg, lab, _, _ = anchor_targets_device(anchors, shapes, annotations_group, num_classes,
                                           negative_overlap, positive_overlap)
    return reg.cpu().numpy(), lab.cpu().numpy()
