"""Host-side image preparation either side of the stem, with the reference's names.

  resize_image / preprocess_image   utils/__init__.py:103-138 (letterbox into a grey 128 square, uint8 kept)
  preprocess_images_device          the same letterbox on the GPU (effdet_letterbox_u8: bit-exact with cv2's 8-bit
                                    bilinear resize), so a raw decoded image goes host -> device once, as bytes
  normalize_image                   utils/__init__.py:87-100, train_tpu.py:130-140,
                                    generators/common.py:418-429 ((v/255 - mean) / std per RGB channel)
  normalization_lut                 the same arithmetic tabulated per byte value: what the device stem
                                    (effdet_stem_conv_u8) applies on the fly, so a uint8 image can be fed to
                                    model.predict_on_batch / train_on_batch directly and the float image
                                    (12 B/pixel) is never built, uploaded or stored.

The reference normalises in float32 with in-place numpy operators (generators/common.py) or TF float32
tensor ops (train_tpu.py); both are: float32(v) / float32(255) -> - float32(mean_c) -> / float32(std_c),
each step rounded to float32.  normalization_lut() evaluates exactly that, so lut[c][v] is bit-identical
to the reference's normalised pixel.
"""
import numpy as np

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def normalize_image(image):
    """uint8 or float (…,3) RGB -> float32, (v/255 - mean)/std in float32 like the reference.  Returns a new
    array (the reference's utils.normalize_image works in place on a float image that was already /255)."""
    x = np.asarray(image).astype(np.float32)
    x /= np.float32(255.0)
    for c in range(3):
        x[..., c] -= np.float32(MEAN[c])
        x[..., c] /= np.float32(STD[c])
    return x


def normalization_lut():
    """(3, 256) float32: lut[c][v] = normalize_image of byte value v in channel c."""
    v = np.arange(256, dtype=np.uint8)
    rgb = np.stack([v, v, v], axis=-1)          # (256, 3)
    return np.ascontiguousarray(normalize_image(rgb).T)


def resize_image(image, image_size):
    """utils/__init__.py:103-132: scale the longer side to image_size (cv2.resize, bilinear), paste into the
    centre of a grey (128) square of the input dtype.  -> (new_image, scale, offset_h, offset_w)."""
    import cv2
    h, w = image.shape[:2]
    if h == w == image_size:
        return image, 0, 0, 0
    if h > w:
        scale = image_size / h
        rh, rw = image_size, int(w * scale)
    else:
        scale = image_size / w
        rh, rw = int(h * scale), image_size
    image = cv2.resize(image, (rw, rh))
    off_h, off_w = (image_size - rh) // 2, (image_size - rw) // 2
    new_image = 128 * np.ones((image_size, image_size, 3), dtype=image.dtype)
    new_image[off_h:off_h + rh, off_w:off_w + rw] = image
    return new_image, scale, off_h, off_w


def preprocess_image(image, image_size):
    """utils/__init__.py:135-138.  The uint8 result can be passed to the model as is."""
    return resize_image(image, image_size)


def preprocess_images_device(images, image_size, out=None, device=None):
    """Letterbox a list of raw uint8 RGB images (each (h, w, 3), numpy or CUDA tensor, any sizes) on the GPU.
    -> (batch (B, image_size, image_size, 3) uint8 CUDA tensor, [(scale, offset_h, offset_w), ...]) with the values
    utils.preprocess_image (utils/__init__.py:135-138) returns per image.  The batch can be passed to
    predict_on_batch / train_on_batch as is (the stem normalises)."""
    import ctypes
    import torch
    from .. import _lib
    lib = _lib.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    B = len(images)
    if out is None:
        out = torch.empty((B, image_size, image_size, 3), dtype=torch.uint8, device=device)
    assert out.shape == (B, image_size, image_size, 3) and out.dtype == torch.uint8 and out.is_cuda
    meta, keep = [], []
    st = _lib.stream_ptr(device)
    for i, img in enumerate(images):
        if not isinstance(img, torch.Tensor):
            a = np.ascontiguousarray(img)
            if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
                raise ValueError("images must be (h, w, 3) uint8, got %s %s" % (a.shape, a.dtype))
            img = torch.from_numpy(a).to(device, non_blocking=True)
        elif img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3 or not img.is_cuda:
            raise ValueError("tensor images must be (h, w, 3) uint8 CUDA tensors")
        img = img.contiguous()
        keep.append(img)
        h, w = int(img.shape[0]), int(img.shape[1])
        rh, rw, oh, ow = (ctypes.c_int() for _ in range(4))
        sc = ctypes.c_double()
        _lib.call("effdet_letterbox_geometry", h, w, image_size, ctypes.byref(rh), ctypes.byref(rw),
                  ctypes.byref(oh), ctypes.byref(ow), ctypes.byref(sc))
        _lib.call("effdet_letterbox_u8", img.data_ptr(), h, w, 0, out[i].data_ptr(), image_size, st)
        s = sc.value
        meta.append((0 if s == 0.0 else s, oh.value, ow.value))
    torch.cuda.current_stream(device).synchronize()      # the staged source tensors may be freed now
    return out, meta
