"""Thin ctypes caller of the plan-level C ABI (csrc/plan.cu, include/effdet_b200.h "plan level"): the whole
inference path of `efficientdet(phi, ...)` / `prediction_model.predict_on_batch` (model.py:356-452,
inference.py:57-59) lowered in C++ -- no torch, only numpy host buffers cross this boundary.  This is the binding a
non-Python host would write in its own FFI; model.Model keeps using engine.Plan (which also serves training and the
per-level taps of the parity tests)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import BF16, F32

U8_INPUT, NO_GRAPH = 1, 2


class CPlan:
    def __init__(self, phi, image_size, batch, num_classes=20, weighted_bifpn=False, dtype="fp32", u8_input=False,
                 graph=True):
        _lib.load()
        self.phi, self.S, self.B, self.C = int(phi), int(image_size), int(batch), int(num_classes)
        self.u8 = bool(u8_input)
        h = ctypes.c_void_p()
        flags = (U8_INPUT if u8_input else 0) | (0 if graph else NO_GRAPH)
        _lib.call("effdet_plan_create", self.phi, self.S, self.B, self.C, int(bool(weighted_bifpn)),
                  {"fp32": F32, "bf16": BF16}[dtype], flags, ctypes.byref(h))
        self._h = h
        self.N = int(_lib.load().effdet_plan_num_anchors(h))

    def close(self):
        if self._h:
            _lib.call("effdet_plan_destroy", self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def weight_manifest(self):
        """[(keras name, shape)] in the creation order of the reference graph."""
        lib = _lib.load()
        out = []
        for i in range(lib.effdet_plan_num_weights(self._h)):
            name, nd, dims = ctypes.c_char_p(), ctypes.c_int(), (ctypes.c_int * 4)()
            _lib.call("effdet_plan_weight_info", self._h, i, ctypes.byref(name), ctypes.byref(nd), dims)
            out.append((name.value.decode(), tuple(dims[:nd.value])))
        return out

    def bind_weights_host(self, weights):
        """weights: {keras name: float32 ndarray} (e.g. utils.hdf5.load_keras_weights of a reference .h5 file)."""
        keep = [(k.encode(), np.ascontiguousarray(v, np.float32)) for k, v in weights.items()]
        names = (ctypes.c_char_p * len(keep))(*[k for k, _ in keep])
        ptrs = (ctypes.c_void_p * len(keep))(*[v.ctypes.data for _, v in keep])
        _lib.call("effdet_plan_bind_weights_host", self._h, names, ptrs, len(keep))

    def forward_device(self, images_ptr, regression_ptr=None, classification_ptr=None, stream=None):
        _lib.call("effdet_forward", self._h, images_ptr, regression_ptr, classification_ptr, stream)

    def detect_host(self, images, anchors=None, score_threshold=0.01, iou_threshold=0.5, max_detections=300):
        """images: (B,S,S,3) float32 (uint8 for a u8_input plan) host array -> [boxes, scores, labels] host arrays
        (prediction_model.predict_on_batch, model.py:443-450)."""
        img = np.ascontiguousarray(images, np.uint8 if self.u8 else np.float32)
        if img.shape != (self.B, self.S, self.S, 3):
            raise ValueError("expected images of shape %s" % ((self.B, self.S, self.S, 3),))
        a = None
        if anchors is not None:
            a = np.ascontiguousarray(anchors, np.float32).reshape(-1, 4)
            if a.shape[0] != self.N:
                raise ValueError("anchors must be (1, %d, 4)" % self.N)
        boxes = np.empty((self.B, max_detections, 4), np.float32)
        scores = np.empty((self.B, max_detections), np.float32)
        labels = np.empty((self.B, max_detections), np.int32)
        _lib.call("effdet_detect_host", self._h, img.ctypes.data, a.ctypes.data if a is not None else None,
                  float(score_threshold), float(iou_threshold), int(max_detections), boxes.ctypes.data,
                  scores.ctypes.data, labels.ctypes.data)
        return [boxes, scores, labels]


class CReplay:
    """ctypes caller of the compiled-plan training API (csrc/replay.cu): effdet_replay_load / region / step.
    Only numpy host arrays cross this boundary (copies through the CUDA runtime, not torch) -- the binding a
    non-Python host would write.  The plan file comes from plan_export.export_train_plan()."""

    def __init__(self, path, graph=True):
        _lib.load()
        h = ctypes.c_void_p()
        _lib.call("effdet_replay_load", str(path).encode(), 0 if graph else 1, ctypes.byref(h))
        self._h = h
        self.num_launches = int(_lib.load().effdet_replay_num_launches(h))

    def close(self):
        if self._h:
            _lib.call("effdet_replay_destroy", self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def region(self, name):
        ptr, n = ctypes.c_void_p(), ctypes.c_size_t()
        _lib.call("effdet_replay_region", self._h, name.encode(), ctypes.byref(ptr), ctypes.byref(n))
        return ptr.value, n.value

    @staticmethod
    def _memcpy(dst, src, nbytes, to_device):
        from cuda.bindings import runtime as cudart
        kind = cudart.cudaMemcpyKind.cudaMemcpyHostToDevice if to_device else cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost
        err, = cudart.cudaMemcpy(dst, src, nbytes, kind)
        if int(err) != 0:
            raise RuntimeError("cudaMemcpy failed: %s" % err)

    def write(self, name, array):
        a = np.ascontiguousarray(array)
        ptr, n = self.region(name)
        if a.nbytes > n:
            raise ValueError("%s: %d bytes into a region of %d" % (name, a.nbytes, n))
        self._memcpy(ptr, a.ctypes.data, a.nbytes, True)

    def read(self, name, dtype, shape=None, offset_bytes=0):
        ptr, n = self.region(name)
        dtype = np.dtype(dtype)
        count = int(np.prod(shape)) if shape is not None else (n - offset_bytes) // dtype.itemsize
        out = np.empty(count, dtype)
        self._memcpy(out.ctypes.data, ptr + offset_bytes, out.nbytes, False)
        return out.reshape(shape) if shape is not None else out

    def step(self, learning_rate, stream=None):
        _lib.call("effdet_replay_step", self._h, float(learning_rate), stream)
