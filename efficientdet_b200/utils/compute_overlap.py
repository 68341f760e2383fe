"""compute_overlap -- mirror of the reference's only native component
(utils/compute_overlap.pyx:13-53, Cython) on effdet_compute_overlap (csrc/targets.cu)."""
import numpy as np
import torch

from .. import _lib
from .._tensor import as_device, give_back


def compute_overlap(boxes, query_boxes):
    """(N,4),(K,4) float64 -> (N,K) float64 IoU with the legacy "+1" pixel convention."""
    if isinstance(boxes, np.ndarray) and boxes.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double' but got '%s'" % boxes.dtype)
    b, hb = as_device(boxes, torch.float64)
    q, hq = as_device(query_boxes, torch.float64)
    if b.dim() != 2 or q.dim() != 2 or b.shape[1] != 4 or q.shape[1] != 4:
        raise ValueError("Buffer has wrong number of dimensions (expected (N,4) and (K,4))")
    out = torch.empty((b.shape[0], q.shape[0]), dtype=torch.float64, device=b.device)
    _lib.call("effdet_compute_overlap", b.data_ptr(), b.shape[0], q.data_ptr(), q.shape[0],
              out.data_ptr(), _lib.stream_ptr())
    return give_back(out, hb and hq)
