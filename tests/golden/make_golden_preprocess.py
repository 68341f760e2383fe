"""Generates tests/golden/preprocess.npz by EXECUTING THE REFERENCE's own utils.resize_image and
utils.normalize_image (/root/reference/utils/__init__.py:87-132; cv2 and numpy are its only imports) on random
uint8 images, following the float path of generators/common.py:406-430 (resize -> float32 grey canvas ->
/255 -> normalize_image in place).  Build container only (needs /root/reference).
Run:  python tests/golden/make_golden_preprocess.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import utils as ref_utils  # noqa: E402  (reference package)

rng = np.random.default_rng(20261018)
out = {}
cases = [(37, 53, 64), (80, 45, 64), (64, 64, 64), (21, 96, 96),
         # larger / up- and down-scaling / already-long-side cases for the letterbox resize itself
         (120, 160, 128), (375, 500, 128), (50, 33, 160), (300, 128, 128), (7, 9, 96), (128, 200, 64)]
out["cases"] = np.array(cases)
for i, (h, w, size) in enumerate(cases):
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    boxed, scale, oh, ow = ref_utils.resize_image(img, size)            # uint8 letterbox, 128 padding
    assert boxed.dtype == np.uint8 and boxed.shape == (size, size, 3)
    f = boxed.astype(np.float32)
    f /= 255.                                                           # generators/common.py:417
    ref_utils.normalize_image(f)                                        # in place, :418-429 == utils/__init__.py:87-100
    out["img_%d" % i] = img
    out["boxed_%d" % i] = boxed
    out["meta_%d" % i] = np.array([scale, oh, ow], np.float64)
    out["norm_%d" % i] = f
# every byte value in every channel
v = np.arange(256, dtype=np.uint8)
ramp = np.stack([v, v, v], -1)[None]
f = ramp.astype(np.float32)
f /= 255.
ref_utils.normalize_image(f)
out["ramp_norm"] = f[0]
np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)
print("wrote", os.path.join(HERE, "preprocess.npz"), {k: v.shape for k, v in out.items()})
