"""CPU: the input normalisation either side of the stem against vectors produced by
EXECUTING the reference's utils.resize_image / utils.normalize_image (tests/golden/make_golden_preprocess.py):
the oracle restatement, the host utilities with the reference's names, and the per-byte table the device stem
applies -- all bit-exact."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "preprocess.npz"))


def test_oracle_normalisation_matches_reference(gold):
    from oracle import preprocess as op
    for i in range(len(gold["cases"])):
        assert np.array_equal(op.normalize_image_ref(gold["boxed_%d" % i]), gold["norm_%d" % i])
    v = np.arange(256, dtype=np.uint8)
    assert np.array_equal(op.normalize_image_ref(np.stack([v, v, v], -1)), gold["ramp_norm"])


def test_host_preprocess_and_lut_bit_exact(gold):
    from efficientdet_b200.utils import preprocess as P
    lut = P.normalization_lut()
    assert lut.shape == (3, 256) and lut.dtype == np.float32
    assert np.array_equal(lut.T, gold["ramp_norm"])
    for i, (h, w, size) in enumerate(gold["cases"]):
        boxed, scale, oh, ow = P.preprocess_image(gold["img_%d" % i], int(size))     # cv2 letterbox, uint8
        assert boxed.dtype == np.uint8
        assert np.array_equal(boxed, gold["boxed_%d" % i])
        assert np.array_equal(np.array([scale, oh, ow], np.float64), gold["meta_%d" % i])
        want = gold["norm_%d" % i]
        assert np.array_equal(P.normalize_image(boxed), want)
        # the table lookup the device stem performs
        got = np.stack([lut[c][boxed[..., c]] for c in range(3)], -1)
        assert np.array_equal(got, want)


def test_oracle_letterbox_resize_matches_reference(gold):
    """oracle.preprocess.resize_image_ref (cv2's 8-bit bilinear resize restated + the reference's letterbox) against
    outputs of the reference's own utils.resize_image (which calls cv2.resize): bit-exact, including the returned
    (scale, offset_h, offset_w)."""
    from oracle import preprocess as op
    for i, (h, w, size) in enumerate(gold["cases"]):
        boxed, scale, oh, ow = op.resize_image_ref(gold["img_%d" % i], int(size))
        assert boxed.dtype == np.uint8
        assert np.array_equal(boxed, gold["boxed_%d" % i]), (i, h, w, size)
        assert np.array_equal(np.array([scale, oh, ow], np.float64), gold["meta_%d" % i])
