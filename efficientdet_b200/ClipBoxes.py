"""ClipBoxes layer -- mirror of the reference's ClipBoxes.py:4-27 on effdet_clip_boxes."""
import torch

from . import _lib
from ._tensor import as_device, give_back
from .keras_compat import Layer


def clip_boxes(image_shape, boxes):
    """image_shape: (B, H, W, C); boxes (B, N, 4): x -> [0, W-1], y -> [0, H-1]."""
    b, host = as_device(boxes)
    if b.dim() != 3 or b.shape[-1] != 4:
        raise ValueError("boxes must be (B, N, 4)")
    out = torch.empty_like(b)
    _lib.call("effdet_clip_boxes", b.data_ptr(), b.shape[0], b.shape[1], float(image_shape[1]),
              float(image_shape[2]), out.data_ptr(), _lib.stream_ptr())
    return give_back(out, host)


class ClipBoxes(Layer):
    """Clips box coordinates to the image tensor's spatial shape."""

    def call(self, inputs, **kwargs):
        image, boxes = inputs
        shape = tuple(image.shape) if hasattr(image, "shape") else tuple(image)
        return clip_boxes(shape, boxes)

    def compute_output_shape(self, input_shape):
        return input_shape[1]
