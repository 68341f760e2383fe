"""Layer API -- mirror of the reference's layers.py: wBiFPNAdd (:11-39) plus the re-exports
RegressBoxes, ClipBoxes, FilterDetections (:6-8)."""
import torch

from . import _lib
from ._lib import BF16, F32
from ._tensor import as_device, device, give_back
from .keras_compat import Layer
from .RegressBoxes import RegressBoxes  # noqa: F401
from .ClipBoxes import ClipBoxes  # noqa: F401
from .FilterDetections import FilterDetections  # noqa: F401


class wBiFPNAdd(Layer):
    """Fast normalised fusion: sum_i relu(w_i) x_i / (sum_i relu(w_i) + epsilon)."""

    def __init__(self, epsilon=1e-4, **kwargs):
        super(wBiFPNAdd, self).__init__(**kwargs)
        self.epsilon = epsilon

    def build(self, input_shape):
        num_in = len(input_shape)
        self.w = self.add_weight(self.name, (num_in,),
                                 torch.full((num_in,), 1.0 / num_in, dtype=torch.float32,
                                            device=device()), trainable=True)

    def call(self, inputs, **kwargs):
        if len(inputs) not in (2, 3):
            raise ValueError("wBiFPNAdd takes 2 or 3 inputs")
        first = inputs[0]
        dt = torch.bfloat16 if isinstance(first, torch.Tensor) and first.dtype == torch.bfloat16 \
            else torch.float32
        ts, hosts = zip(*[as_device(x, dt) for x in inputs])
        for t in ts[1:]:
            if t.shape != ts[0].shape:
                raise ValueError("wBiFPNAdd inputs must have the same shape")
        out = torch.empty_like(ts[0])
        ptrs = (_lib.c_void_p * 3)(*[t.data_ptr() for t in ts], *([None] * (3 - len(ts))))
        _lib.call("effdet_wbifpn_add", ptrs, len(ts), self.w.data_ptr(), float(self.epsilon),
                  out.data_ptr(), out.numel(), BF16 if dt == torch.bfloat16 else F32,
                  _lib.stream_ptr())
        return give_back(out, all(hosts))

    def compute_output_shape(self, input_shape):
        return input_shape[0]

    def get_config(self):
        config = super(wBiFPNAdd, self).get_config()
        config.update({"epsilon": self.epsilon})
        return config
