"""Pins the oracle (numpy + C restatements) to vectors produced by executing the
reference's utils/anchors.py and Cython compute_overlap (tests/golden/make_golden.py),
and to SURVEY.md Appendix B's known answers."""
import hashlib

import numpy as np

from oracle import anchors as oa
from oracle import overlap_c, build_ref, graph


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_generate_anchors_bit_exact(golden):
    for s in (16, 32, 64, 128, 256, 512, 48):
        assert np.array_equal(oa.generate_anchors(s), golden["gen_%d" % s]), s


def test_generate_anchors_kat():
    a = oa.generate_anchors(32)
    np.testing.assert_allclose(a[0], [-22.627416998, -11.313708499, 22.627416998, 11.313708499],
                               rtol=0, atol=1e-9)
    np.testing.assert_allclose(a[3], [-16, -16, 16, 16], rtol=0, atol=0)
    np.testing.assert_allclose(a[4][0], -20.1587371826, rtol=0, atol=1e-9)


def test_anchors_for_small_shapes_bit_exact(golden):
    for shp in ((128, 128), (96, 160), (100, 150)):
        assert np.array_equal(oa.anchors_for_shape(shp), golden["anchors_%dx%d" % shp]), shp


def test_anchors_for_model_sizes_digest(golden):
    for i, S in enumerate(golden["model_sizes"]):
        a = oa.anchors_for_shape((int(S), int(S)))
        assert a.dtype == np.float64
        assert a.shape[0] == golden["model_counts"][i]
        assert a.sum() == golden["model_sums"][i]
        assert sha(a) == str(golden["model_sha_f64"][i])
        assert sha(a.astype(np.float32)) == str(golden["model_sha_f32"][i])
        assert np.array_equal(a[0], golden["anchors_%d_first" % S])
        assert np.array_equal(a[-1], golden["anchors_%d_last" % S])


def test_appendix_b_counts():
    for S, n, tot in ((512, 49104, 50282496), (768, 110484, 169703424), (1024, 196416, 402259968),
                      (1280, 306900, 785664000), (1408, 371349, 1045718784)):
        a = oa.anchors_for_shape((S, S))
        assert a.shape == (n, 4) and a.sum() == tot
        np.testing.assert_allclose(a[0], [-18.627416998, -7.313708499, 26.627416998, 15.313708499],
                                   atol=1e-9)


def test_compute_overlap_numpy_and_c_bit_exact(golden):
    b, q, want = golden["ov_boxes"], golden["ov_query"], golden["ov_result"]
    assert np.array_equal(oa.compute_overlap(b, q), want)
    assert np.array_equal(overlap_c.compute_overlap(b, q), want)
    kat = oa.compute_overlap([[0, 0, 10, 10], [5, 5, 15, 15]], [[0, 0, 10, 10]])
    assert np.array_equal(kat, golden["ov_kat"])
    assert kat[0, 0] == 1.0 and abs(kat[1, 0] - 0.17475728) < 1e-8


def test_compute_overlap_against_compiled_reference(golden):
    ref = build_ref.load_ref_compute_overlap()
    if ref is None:
        import pytest
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(5)
    xy = rng.uniform(0, 300, (500, 2)); wh = rng.uniform(0.5, 150, (500, 2))
    b = np.concatenate([xy, xy + wh], 1)
    q = b[rng.integers(0, 500, 12)] + rng.normal(0, 3, (12, 4))
    assert np.array_equal(ref(b, q), overlap_c.compute_overlap(b, q))
    assert np.array_equal(ref(b, q), oa.compute_overlap(b, q))


def test_bbox_transform_bit_exact(golden):
    a = oa.anchors_for_shape((128, 128))
    assert np.array_equal(oa.bbox_transform(a, golden["bt_gt"]), golden["bt_result"])


def _kat_inputs():
    return ([(512, 512, 3)],
            [{"bboxes": np.array([[100, 120, 300, 360], [10, 10, 60, 80]], np.float32),
              "labels": np.array([3, 7], np.float32)}])


def test_anchor_targets_kat(golden):
    a = oa.anchors_for_shape((512, 512))
    shapes, ann = _kat_inputs()
    for impl in (oa.anchor_targets_bbox, overlap_c.anchor_targets_bbox):
        reg, lab = impl(a, shapes, ann, 20)
        assert reg.shape == (1, 49104, 5) and lab.shape == (1, 49104, 21)
        pos = np.nonzero(reg[0, :, 4] == 1)[0]
        assert len(pos) == 68 and (reg[0, :, 4] == -1).sum() == 164
        assert list(pos[:12]) == [1769, 2336, 2339, 2345, 2348, 2354, 2357, 2912, 2915, 2921, 2924,
                                  2930]
        np.testing.assert_allclose(reg[0, 1769], [-0.059214663, 0.728236, -0.137648, 2.6184294, 1],
                                   rtol=1e-6)
        assert np.array_equal(pos, golden["kat_pos_idx"])
        assert np.array_equal(np.nonzero(reg[0, :, 4] == -1)[0], golden["kat_ign_idx"])
        assert sha(reg) == str(golden["kat_reg_sha"])
        assert sha(lab) == str(golden["kat_lab_sha"])


def test_anchor_targets_ragged_batch_bit_exact(golden):
    a = oa.anchors_for_shape((128, 128))
    shapes = [tuple(s) for s in golden["tg_img_shapes"]]
    ann = [{"bboxes": golden["tg_bboxes_%d" % i], "labels": golden["tg_labels_%d" % i]}
           for i in range(len(shapes))]
    for impl in (oa.anchor_targets_bbox, overlap_c.anchor_targets_bbox):
        reg, lab = impl(a, shapes, ann, 6)
        assert np.array_equal(reg, golden["tg_regression"])
        assert np.array_equal(lab, golden["tg_labels"])


def test_backbone_layer_counts():
    # train_tpu.py:24 EFFICIENTNET_DEPTHS -- validates block/repeat/skip/SE structure
    assert [graph.keras_layer_count(p) for p in range(7)] == [227, 329, 329, 374, 464, 566, 656]
