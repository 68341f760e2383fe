"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's input
normalisation, the step just upstream of the stem.

  normalize_image_ref   generators/common.py:418-429 and train_tpu.py:130-140 (identical arithmetic):
                        float32 image / 255., then per channel  -= mean[c], /= std[c]  (numpy in-place
                        float32 operators; python floats are cast to float32 by the in-place op)
  letterbox_ref         generators/common.py:406-417 without the cv2.resize (the caller resizes): float32 grey
                        (128.) canvas with the image pasted in the centre

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
