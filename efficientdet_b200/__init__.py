"""efficientdet_b200 -- B200-native EfficientDet hot path behind the reference's Python API.

Host side: thin Python mirroring Ely-S/EfficientDet's builder and layer interface
(model.efficientdet, layers.wBiFPNAdd / RegressBoxes / ClipBoxes / FilterDetections,
utils.anchors).  Device side: hand-written sm_100a CUDA kernels in
libeffdet_b200.so, bound through the C ABI of include/effdet_b200.h.
torch is used for device memory, streams/graphs and torch.distributed only.
There is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
