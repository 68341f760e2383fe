"""numpy stand-in for the slice of `tensorflow` / `tensorflow.keras` that the reference's FilterDetections.py uses (and
RegressBoxes.py / ClipBoxes.py: keras.backend.stack, tf.clip_by_value), so that the reference files can be imported
and EXECUTED unmodified in the build container (TensorFlow is not
installable offline; requirements.txt:21).  Test infrastructure only: tests/golden/make_golden_filter.py is its one
user.  Written independently of oracle/tail.py -- the fixtures it produces are what pins that oracle.

What is executed from the reference: the whole control flow of filter_by_score_and_nms / filter_detections / the
FilterDetections layer (FilterDetections.py:5-190): thresholding, the per-class loop, label construction, index
gathering, concatenation order, top-k selection, gathers, -1 padding, dtype casts, `nms=False -> iou_threshold = 0`,
tf.map_fn over the batch.  What is RESTATED here from TensorFlow's documented behaviour:
  * tf.image.non_max_suppression (NonMaxSuppressionV3): candidates in descending score order; a candidate is
    dropped when its IoU with an already selected box is > iou_threshold; stops at max_output_size; boxes may have
    their corners in any order (min / max per axis); IoU is 0 when either area is <= 0 or the intersection is empty;
    float32 arithmetic.
  * tf.nn.top_k: values in descending order, equal values keep the lower index first.
  * tf.compat.v1.where(cond): coordinates of the true elements in row-major order, int64.
"""
import builtins
import sys
import types

import numpy as np

F = np.float32


class T(np.ndarray):
    """ndarray with the two Tensor methods the reference calls."""

    def set_shape(self, shape):
        assert tuple(self.shape) == tuple(int(s) for s in shape), (self.shape, shape)


def t(a, dtype=None):
    return np.asarray(a, dtype=dtype).view(T)


def _where(cond, x=None, y=None, name=None):
    assert x is None and y is None
    return t(np.argwhere(np.asarray(cond)).astype(np.int64))


def _gather_nd(params, indices, name=None):
    params, indices = np.asarray(params), np.asarray(indices)
    return t(params[tuple(indices[:, k] for k in range(indices.shape[1]))])


def _gather(params, indices, axis=0, name=None):
    return t(np.take(np.asarray(params), np.asarray(indices), axis=0))


def _iou(a, b):
    ay0, ay1 = min(a[0], a[2]), max(a[0], a[2])
    ax0, ax1 = min(a[1], a[3]), max(a[1], a[3])
    by0, by1 = min(b[0], b[2]), max(b[0], b[2])
    bx0, bx1 = min(b[1], b[3]), max(b[1], b[3])
    area_a = F(ay1 - ay0) * F(ax1 - ax0)
    area_b = F(by1 - by0) * F(bx1 - bx0)
    if area_a <= 0 or area_b <= 0:
        return F(0)
    iy0, ix0 = max(ay0, by0), max(ax0, bx0)
    iy1, ix1 = min(ay1, by1), min(ax1, bx1)
    inter = F(max(F(iy1 - iy0), F(0))) * F(max(F(ix1 - ix0), F(0)))
    return F(inter / F(F(area_a + area_b) - inter))


def _non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5, score_threshold=float("-inf"), name=None):
    boxes = np.asarray(boxes, F).reshape(-1, 4)
    scores = np.asarray(scores, F)
    order = sorted(range(scores.shape[0]), key=lambda i: (-float(scores[i]), i))
    thr = F(iou_threshold)
    keep = []
    for i in order:
        if len(keep) >= int(max_output_size):
            break
        if all(not (_iou(boxes[i], boxes[j]) > thr) for j in keep):
            keep.append(i)
    return t(np.asarray(keep, np.int32))


def _top_k(values, k=1, sorted=True, name=None):
    values = np.asarray(values)
    order = np.asarray(builtins.sorted(range(values.shape[0]), key=lambda i: (-float(values[i]), i))[:int(k)], np.int32)
    return t(values[order]), t(order)


def _pad(tensor, paddings, mode="CONSTANT", constant_values=0, name=None):
    tensor = np.asarray(tensor)
    pads = [(int(a), int(b)) for a, b in paddings]
    return t(np.pad(tensor, pads, mode="constant", constant_values=constant_values))


def _map_fn(fn, elems, dtype=None, parallel_iterations=None, **kw):
    n = int(np.asarray(elems[0]).shape[0])
    outs = [fn([t(np.asarray(e)[i]) for e in elems]) for i in range(n)]
    return [t(np.stack([np.asarray(o[j]) for o in outs]).astype(dt)) for j, dt in enumerate(dtype)]


class Layer:
    def __init__(self, name=None, **kwargs):
        self.name = name

    def __call__(self, inputs, **kwargs):
        return self.call(inputs, **kwargs)

    def get_config(self):
        return {"name": self.name}


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


backend = _module(
    "tensorflow.keras.backend", floatx=lambda: "float32",
    greater=lambda a, b: t(np.asarray(a) > b),
    gather=_gather,
    shape=lambda x: np.asarray(np.asarray(x).shape, np.int64),
    concatenate=lambda xs, axis=-1: t(np.concatenate([np.asarray(x) for x in xs], axis=axis)),
    max=lambda x, axis=None: t(np.asarray(x).max(axis=axis)),
    argmax=lambda x, axis=-1: t(np.asarray(x).argmax(axis=axis).astype(np.int64)),
    minimum=lambda a, b: np.minimum(a, b), maximum=lambda a, b: np.maximum(a, b),
    cast=lambda x, dtype: t(np.asarray(x).astype(dtype)),
    stack=lambda xs, axis=0: t(np.stack([np.asarray(x) for x in xs], axis=axis)))
layers = _module("tensorflow.keras.layers", Layer=Layer)
keras = _module("tensorflow.keras", backend=backend, layers=layers)
image = _module("tensorflow.image", non_max_suppression=_non_max_suppression)
nn = _module("tensorflow.nn", top_k=_top_k)
compat_v1 = _module("tensorflow.compat.v1", where=_where)
compat = _module("tensorflow.compat", v1=compat_v1)
tf = _module("tensorflow", keras=keras, image=image, nn=nn, compat=compat, gather_nd=_gather_nd, gather=_gather,
             stack=lambda xs, axis=0: t(np.stack([np.asarray(x) for x in xs], axis=axis)),
             ones=lambda shape, dtype="float32": t(np.ones(tuple(int(s) for s in shape), dtype=dtype)),
             pad=_pad, map_fn=_map_fn,
             clip_by_value=lambda x, lo, hi, name=None: t(np.minimum(np.maximum(np.asarray(x), F(lo)), F(hi))))


def install():
    sys.modules.update({"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.backend": backend,
                        "tensorflow.keras.layers": layers, "tensorflow.image": image, "tensorflow.nn": nn,
                        "tensorflow.compat": compat, "tensorflow.compat.v1": compat_v1})
