"""Oracle: anchor table, IoU, anchor targets (numpy, float64 like the reference).

Restates /root/reference/utils/anchors.py:
  AnchorParameters.default :46-52   generate_anchors :372-403   shift :339-369
  guess_shapes :280-293             anchors_for_shape :296-336
  compute_gt_annotations :210-239   anchor_targets_bbox :130-207
  bbox_transform :406-439
and /root/reference/utils/compute_overlap.pyx:13-53 (the "+1" pixel IoU).
TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import numpy as np

SIZES = [32, 64, 128, 256, 512]
STRIDES = [8, 16, 32, 64, 128]
# keras.backend.floatx() == 'float32' in the reference => ratios/scales are
# rounded to float32 first and then promoted to float64 in the arithmetic.
RATIOS = np.array([0.5, 1, 2], np.float32)
SCALES = np.array([2 ** 0, 2 ** (1.0 / 3.0), 2 ** (2.0 / 3.0)], np.float32)


def generate_anchors(base_size=16, ratios=None, scales=None):
    ratios = RATIOS if ratios is None else ratios
    scales = SCALES if scales is None else scales
    nr, ns = len(ratios), len(scales)
    out = np.zeros((nr * ns, 4))
    for r in range(nr):
        for s in range(ns):
            # base_size * np.tile(scales, ...) is evaluated in the dtype of
            # `scales` (float32) before being stored into the float64 table
            # (utils/anchors.py:391).
            side = np.float64(np.asarray(scales)[s] * base_size)
            area = side * side
            w = np.sqrt(area / np.float64(ratios[r]))
            h = w * np.float64(ratios[r])
            out[r * ns + s] = (0 - w * 0.5, 0 - h * 0.5, w - w * 0.5, h - h * 0.5)
    return out


def guess_shapes(image_shape, pyramid_levels):
    shp = np.array(image_shape[:2])
    return [(shp + 2 ** x - 1) // (2 ** x) for x in pyramid_levels]


def shift(shape, stride, anchors):
    sx = (np.arange(0, shape[1]) + 0.5) * stride
    sy = (np.arange(0, shape[0]) + 0.5) * stride
    A = anchors.shape[0]
    out = np.empty((shape[0], shape[1], A, 4))
    out[..., 0] = anchors[None, None, :, 0] + sx[None, :, None]
    out[..., 1] = anchors[None, None, :, 1] + sy[:, None, None]
    out[..., 2] = anchors[None, None, :, 2] + sx[None, :, None]
    out[..., 3] = anchors[None, None, :, 3] + sy[:, None, None]
    return out.reshape(-1, 4)


def anchors_for_shape(image_shape, pyramid_levels=None, sizes=None, strides=None,
                      ratios=None, scales=None):
    pyramid_levels = [3, 4, 5, 6, 7] if pyramid_levels is None else pyramid_levels
    sizes = SIZES if sizes is None else sizes
    strides = STRIDES if strides is None else strides
    shapes = guess_shapes(image_shape, pyramid_levels)
    parts = []
    for i, _ in enumerate(pyramid_levels):
        a = generate_anchors(sizes[i], ratios, scales)
        parts.append(shift(shapes[i], strides[i], a))
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, 4))


def compute_overlap(boxes, query):
    """(N,4),(K,4) float64 -> (N,K) IoU with the +1 convention; same op order
    as compute_overlap.pyx:31-52 so results are bit-identical."""
    boxes = np.asarray(boxes, np.float64)
    query = np.asarray(query, np.float64)
    N, K = boxes.shape[0], query.shape[0]
    out = np.zeros((N, K))
    for k in range(K):
        qa = (query[k, 2] - query[k, 0] + 1) * (query[k, 3] - query[k, 1] + 1)
        iw = np.minimum(boxes[:, 2], query[k, 2]) - np.maximum(boxes[:, 0], query[k, 0]) + 1
        ih = np.minimum(boxes[:, 3], query[k, 3]) - np.maximum(boxes[:, 1], query[k, 1]) + 1
        ok = (iw > 0) & (ih > 0)
        ua = (boxes[:, 2] - boxes[:, 0] + 1) * (boxes[:, 3] - boxes[:, 1] + 1) + qa - iw * ih
        with np.errstate(divide="ignore", invalid="ignore"):
            v = iw * ih / ua
        out[ok, k] = v[ok]
    return out


def bbox_transform(anchors, gt, mean=None, std=None):
    mean = np.array([0, 0, 0, 0]) if mean is None else np.asarray(mean)
    std = np.array([0.2, 0.2, 0.2, 0.2]) if std is None else np.asarray(std)
    aw = anchors[:, 2] - anchors[:, 0]
    ah = anchors[:, 3] - anchors[:, 1]
    t = np.stack(((gt[:, 0] - anchors[:, 0]) / aw, (gt[:, 1] - anchors[:, 1]) / ah,
                  (gt[:, 2] - anchors[:, 2]) / aw, (gt[:, 3] - anchors[:, 3]) / ah)).T
    return (t - mean) / std


def compute_gt_annotations(anchors, annotations, negative_overlap=0.4, positive_overlap=0.5):
    ov = compute_overlap(anchors.astype(np.float64), annotations.astype(np.float64))
    arg = np.argmax(ov, axis=1)
    mx = ov[np.arange(ov.shape[0]), arg]
    pos = mx >= positive_overlap
    ign = (mx > negative_overlap) & ~pos
    return pos, ign, arg


def anchor_targets_bbox(anchors, image_shapes, annotations_group, num_classes,
                        negative_overlap=0.4, positive_overlap=0.5):
    """image_shapes: list of (H, W[, C]) tuples (the reference reads image.shape).
    Returns (regression (B,N,5) f32, labels (B,N,C+1) f32)."""
    assert len(image_shapes) == len(annotations_group)
    assert len(annotations_group) > 0
    B, N = len(image_shapes), anchors.shape[0]
    reg = np.zeros((B, N, 5), np.float32)
    lab = np.zeros((B, N, num_classes + 1), np.float32)
    for i, (shp, ann) in enumerate(zip(image_shapes, annotations_group)):
        bb = np.asarray(ann["bboxes"])
        if bb.shape[0]:
            pos, ign, arg = compute_gt_annotations(anchors, bb, negative_overlap, positive_overlap)
            lab[i, ign, -1] = -1
            lab[i, pos, -1] = 1
            reg[i, ign, -1] = -1
            reg[i, pos, -1] = 1
            lab[i, pos, np.asarray(ann["labels"])[arg[pos]].astype(int)] = 1
            reg[i, :, :-1] = bbox_transform(anchors, bb[arg, :])
        if len(shp):
            cx = (anchors[:, 0] + anchors[:, 2]) / 2
            cy = (anchors[:, 1] + anchors[:, 3]) / 2
            out = np.logical_or(cx >= shp[1], cy >= shp[0])
            lab[i, out, -1] = -1
            reg[i, out, -1] = -1
    return reg, lab
