"""RegressBoxes layer + apply_bbox_deltas -- mirror of the reference's RegressBoxes.py
(:13-123 layer, :126-164 function) running on effdet_regress_boxes (csrc/tail.cu)."""
import numpy as np
import torch

from . import _lib
from ._tensor import as_device, give_back
from .keras_compat import Layer

default_mean = np.array([0, 0, 0, 0], dtype="float32")
default_std = np.array([0.2, 0.2, 0.2, 0.2], dtype="float32")
default_mean.setflags(write=False)
default_std.setflags(write=False)


def _f4(v):
    v = np.asarray(v, np.float32).reshape(-1)
    if v.shape[0] != 4:
        raise ValueError("mean/std must have 4 entries")
    return (_lib.c_float * 4)(*[float(x) for x in v])


def apply_bbox_deltas(boxes, deltas, mean=default_mean, std=default_std):
    """boxes (B or 1, N, 4), deltas (B, N, 4) -> (B, N, 4); numpy in -> numpy out."""
    d, host_d = as_device(deltas)
    a, host_a = as_device(boxes)
    if d.dim() != 3 or d.shape[-1] != 4 or a.dim() != 3 or a.shape[-1] != 4:
        raise ValueError("boxes and deltas must be (B, N, 4)")
    B, N = d.shape[0], d.shape[1]
    if a.shape[1] != N or a.shape[0] not in (1, B):
        raise ValueError("boxes %s incompatible with deltas %s" % (tuple(a.shape), tuple(d.shape)))
    out = torch.empty_like(d)
    _lib.call("effdet_regress_boxes", a.data_ptr(), int(a.shape[0] == B), d.data_ptr(), _f4(mean), _f4(std), B, N, out.data_ptr(), _lib.stream_ptr())
    return give_back(out, host_d)


class RegressBoxes(Layer):
    """Keras-style layer applying regression deltas to anchor boxes."""

    dtype = "float32"
    saved_anchors = False

    def __init__(self, mean=default_mean, std=default_std, anchor_shape=None, *args, **kwargs):
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
        if isinstance(anchor_shape, (list, tuple)):
            anchor_shape = np.array(anchor_shape)
        elif anchor_shape is not None and not isinstance(anchor_shape, np.ndarray):
            raise ValueError("Expected anchor_shape to be a np.ndarray, list or tuple. Received: "
                             "{}".format(type(anchor_shape)))
        self.mean = mean
        self.std = std
        self.anchor_shape = anchor_shape
        super(RegressBoxes, self).__init__(*args, **kwargs)
        if anchor_shape is not None:
            from ._tensor import device
            self.anchors = self.add_weight(
                "anchor_boxes_baked", tuple(int(s) for s in anchor_shape),
                torch.ones(tuple(int(s) for s in anchor_shape), dtype=torch.float32,
                           device=device()), trainable=False)

    def call(self, inputs, **kwargs):
        if self.saved_anchors:
            anchors, deltas = self.anchors, inputs[0]
        else:
            anchors, deltas = inputs
        return apply_bbox_deltas(anchors, deltas, mean=self.mean, std=self.std)

    def compute_output_shape(self, input_shape):
        if self.saved_anchors:
            return tuple(self.anchors.shape)
        elif len(input_shape) == 2:
            return input_shape[0]
        raise ValueError(self.__class__, "needs to either get anchors as an input"
                                         "or set before use with set_anchors()")

    def set_anchors(self, anchors):
        self.saved_anchors = True
        self.set_weights([np.asarray(anchors).astype("float32")])

    def get_config(self):
        config = super(RegressBoxes, self).get_config()
        config.update({"mean": self.mean.tolist(), "std": self.std.tolist(),
                       "anchor_shape": None if self.anchor_shape is None
                       else self.anchor_shape.tolist()})
        return config
