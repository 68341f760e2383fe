"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's input
normalisation, the step just upstream of the stem.

  normalize_image_ref   generators/common.py:418-429 and train_tpu.py:130-140 (identical arithmetic):
                        float32 image / 255., then per channel  -= mean[c], /= std[c]  (numpy in-place
                        float32 operators; python floats are cast to float32 by the in-place op)
  letterbox_ref         generators/common.py:406-417 without the cv2.resize (the caller resizes): float32 grey
                        (128.) canvas with the image pasted in the centre

  resize_linear_u8_ref  cv2.resize(img, (w, h)) for uint8 images, the call inside utils/__init__.py:122 and
                        generators/common.py:416.  OpenCV is an un-vendored dependency (requirements.txt:13
                        opencv-python==3.4.2.17); this restates its published 8-bit INTER_LINEAR algorithm
                        (modules/imgproc/src/resize.cpp: resizeGeneric_ with HResizeLinear / VResizeLinear<uchar,
                        int, short, FixedPtCast<.., INTER_RESIZE_COEF_BITS * 2>>): 11-bit fixed-point weights.
  resize_image_ref      utils/__init__.py:103-132 on top of it (uint8 letterbox, 128 padding)

PINNED: tests/golden/preprocess.npz, written by tests/golden/make_golden_preprocess.py, which EXECUTES the
reference's own utils.resize_image / utils.normalize_image (utils/__init__.py:87-132) on random uint8 images.
"""
import numpy as np


def normalize_image_ref(image_u8):
    new_image = np.asarray(image_u8).astype(np.float32)
    new_image /= 255.
    mean = [0.485, 0.456, 0.406]
    std = [0.229, 0.224, 0.225]
    for c in range(3):
        new_image[..., c] -= mean[c]
    for c in range(3):
        new_image[..., c] /= std[c]
    return new_image


def letterbox_ref(resized_u8, image_size):
    rh, rw = resized_u8.shape[:2]
    new_image = np.ones((image_size, image_size, 3), dtype=np.float32) * 128.
    oh, ow = (image_size - rh) // 2, (image_size - rw) // 2
    new_image[oh:oh + rh, ow:ow + rw] = resized_u8.astype(np.float32)
    return new_image, oh, ow


def _lin_coeffs(ssize, dsize, clamp_coord):
    scale = 1.0 / (dsize / ssize)                       # inv_scale = dsize / ssize; scale = 1 / inv_scale (double)
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)    # fx = (float)((dx + 0.5) * scale_x - 0.5)
    s = np.floor(f).astype(np.int64)                    # sx = cvFloor(fx)
    f = (f - s.astype(np.float32)).astype(np.float32)   # fx -= sx
    if clamp_coord:                                     # x only: fx = 0 at the borders
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= ssize - 1
        f[hi] = 0
        s[hi] = ssize - 1
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int32)      # saturate_cast<short>(.. * 2048)
    a1 = np.rint(f * np.float32(2048.0)).astype(np.int32)
    return s, a0, a1


def resize_linear_u8_ref(img, dw, dh):
    sh, sw = img.shape[:2]
    if (sh, sw) == (dh, dw):
        return img.copy()
    xi, xa0, xa1 = _lin_coeffs(sw, dw, True)
    yi, ya0, ya1 = _lin_coeffs(sh, dh, False)
    x1 = np.minimum(xi + 1, sw - 1)
    src = img.astype(np.int32)
    H = src[:, xi, :] * xa0[None, :, None] + src[:, x1, :] * xa1[None, :, None]      # horizontal pass, int32
    y0, y1 = np.clip(yi, 0, sh - 1), np.clip(yi + 1, 0, sh - 1)                      # y: rows are clipped instead
    S0, S1 = H[y0], H[y1]
    b0, b1 = ya0[:, None, None], ya1[:, None, None]
    out = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_image_ref(image, image_size):
    h, w = image.shape[:2]
    if h == w == image_size:
        return image, 0, 0, 0
    if h > w:
        scale = image_size / h
        rh, rw = image_size, int(w * scale)
    else:
        scale = image_size / w
        rh, rw = int(h * scale), image_size
    image = resize_linear_u8_ref(image, rw, rh)
    oh, ow = (image_size - rh) // 2, (image_size - rw) // 2
    new_image = 128 * np.ones((image_size, image_size, 3), dtype=image.dtype)
    new_image[oh:oh + rh, ow:ow + rw] = image
    return new_image, scale, oh, ow
