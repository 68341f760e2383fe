"""GPU parity of the network forward kernels against the torch-CPU oracle.
Tolerances (BASELINE.json north_star): <= 1e-4 relative in fp32 mode, <= 2e-2 in bf16 mode,
relative = max|got - want| / max|want| per tensor."""
import numpy as np
import pytest
import torch

from util_model import perturb_weights, rel_err

pytestmark = pytest.mark.gpu

FP32_TOL, BF16_TOL = 1e-4, 2e-2


def _dev(a, dt=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().numpy()


def _nchw(a):
    return torch.from_numpy(a).permute(0, 3, 1, 2).contiguous()


# ------------------------------------------------------------------ single kernels
@pytest.mark.parametrize("k,stride,H,cin,cout,act", [
    (1, 1, 16, 24, 144, 2), (1, 1, 8, 1152, 320, 0), (3, 1, 12, 64, 36, 0), (3, 2, 16, 320, 64, 1),
    (3, 2, 8, 64, 64, 1), (3, 1, 4, 88, 810, 3), (1, 1, 9, 40, 240, 2)])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_conv2d(k, stride, H, cin, cout, act, dtype):
    from efficientdet_b200 import _lib
    from oracle import graph
    import ctypes
    rng = np.random.default_rng(k * 100 + cin)
    B = 3
    tdt = torch.float32 if dtype == "fp32" else torch.bfloat16
    x = rng.standard_normal((B, H, H, cin)).astype(np.float32)
    w = (rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    sh = rng.normal(0, 0.2, cout).astype(np.float32)
    xd = _dev(x, tdt)
    xq = xd.float().cpu().numpy()
    Ho = (H + stride - 1) // stride
    use_gate = k == 1
    gate = rng.uniform(0.2, 1.0, (B, cin)).astype(np.float32)
    res = rng.standard_normal((B, Ho, Ho, cout)).astype(np.float32)
    keep = rng.uniform(0.5, 1.5, B).astype(np.float32)
    resd = _dev(res, tdt)
    out = torch.empty((B, Ho, Ho, cout), dtype=tdt, device="cuda")
    wd, scd, shd, gd, kd = _dev(w), _dev(sc), _dev(sh), _dev(gate), _dev(keep)
    d = _lib.ConvDesc()
    d.n_groups = 1
    d.x[0], d.y[0], d.residual[0] = xd.data_ptr(), out.data_ptr(), resd.data_ptr()
    d.H[0], d.W[0] = H, H
    d.B, d.Cin, d.Cout, d.kh, d.kw, d.stride = B, cin, cout, k, k, stride
    d.weight, d.scale, d.shift = wd.data_ptr(), scd.data_ptr(), shd.data_ptr()
    d.gate = gd.data_ptr() if use_gate else None
    d.keep = kd.data_ptr()
    d.act = act
    d.in_dtype = d.out_dtype = _lib.F32 if dtype == "fp32" else _lib.BF16
    d.allow_tensor_core = 0
    _lib.call("effdet_conv2d", ctypes.byref(d), _lib.stream_ptr())
    torch.cuda.synchronize()
    xin = torch.from_numpy(xq).double()
    if use_gate:
        xin = xin * torch.from_numpy(gate).double()[:, None, None, :]
    y = graph.conv2d(xin.permute(0, 3, 1, 2), w.astype(np.float64), stride)
    y = y * torch.from_numpy(sc).double().view(1, -1, 1, 1) + torch.from_numpy(sh).double().view(1, -1, 1, 1)
    y = [lambda v: v, torch.relu, graph.swish, torch.sigmoid][act](y)
    y = y * torch.from_numpy(keep).double().view(-1, 1, 1, 1) + _nchw(resd.float().cpu().numpy()).double()
    err = rel_err(out.float().cpu().numpy(), _nhwc(y))
    assert err < (1e-5 if dtype == "fp32" else 6e-3), err


@pytest.mark.parametrize("k,stride,H,C", [(3, 1, 16, 32), (3, 2, 16, 96), (5, 2, 12, 144), (5, 1, 7, 240),
                                          (3, 1, 5, 1152)])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_dwconv_and_se(k, stride, H, C, dtype):
    from efficientdet_b200 import _lib
    from oracle import graph
    rng = np.random.default_rng(k * 10 + C)
    B = 2
    tdt = torch.float32 if dtype == "fp32" else torch.bfloat16
    dt = _lib.F32 if dtype == "fp32" else _lib.BF16
    x = rng.standard_normal((B, H, H, C)).astype(np.float32)
    w = (rng.standard_normal((k, k, C, 1)) / k).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, C).astype(np.float32)
    sh = rng.normal(0, 0.2, C).astype(np.float32)
    xd = _dev(x, tdt)
    Ho = (H + stride - 1) // stride
    out = torch.empty((B, Ho, Ho, C), dtype=tdt, device="cuda")
    nblk = _lib.load().effdet_dwconv_se_blocks(B, H, H, C, stride, dt)
    assert nblk >= 1
    se = torch.full((B, nblk, C), float("nan"), dtype=torch.float32, device="cuda")
    wd, scd, shd = _dev(w), _dev(sc), _dev(sh)
    _lib.call("effdet_dwconv", xd.data_ptr(), wd.data_ptr(), scd.data_ptr(), shd.data_ptr(),
              out.data_ptr(), se.data_ptr(), nblk, B, H, H, C, k, stride, _lib.ACT_SWISH, dt,
              _lib.stream_ptr())
    xin = _nchw(xd.float().cpu().numpy()).double()
    y = graph.dwconv2d(xin, w.astype(np.float64), stride)
    y = graph.swish(y * torch.from_numpy(sc).double().view(1, -1, 1, 1) + torch.from_numpy(sh).double().view(1, -1, 1, 1))
    tol = 1e-5 if dtype == "fp32" else 6e-3
    assert rel_err(out.float().cpu().numpy(), _nhwc(y)) < tol
    want_sum = y.sum(dim=(2, 3)).numpy()
    assert rel_err(se.sum(dim=1).cpu().numpy(), want_sum) < 1e-4
    # SE FCs
    R = max(1, C // 24)
    w1 = (rng.standard_normal((C, R)) / np.sqrt(C)).astype(np.float32)
    b1 = rng.normal(0, 0.1, R).astype(np.float32)
    w2 = (rng.standard_normal((R, C)) / np.sqrt(R)).astype(np.float32)
    b2 = rng.normal(0, 0.1, C).astype(np.float32)
    gate = torch.empty((B, C), dtype=torch.float32, device="cuda")
    w1d, b1d, w2d, b2d = _dev(w1), _dev(b1), _dev(w2), _dev(b2)
    _lib.call("effdet_se_gate", se.data_ptr(), nblk, 1.0 / (Ho * Ho), w1d.data_ptr(), b1d.data_ptr(),
              w2d.data_ptr(), b2d.data_ptr(), gate.data_ptr(), B, C, R, _lib.stream_ptr())
    mean = se.sum(dim=1).cpu().double().numpy() / (Ho * Ho)
    r = mean @ w1.astype(np.float64) + b1
    r = r / (1 + np.exp(-r))
    g = 1 / (1 + np.exp(-(r @ w2.astype(np.float64) + b2)))
    assert rel_err(gate.cpu().numpy(), g) < 1e-5


@pytest.mark.parametrize("mode,three,weighted", [(1, False, True), (2, True, True), (2, False, False),
                                                 (1, False, False), (2, True, False), (0, True, True)])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_bifpn_node(mode, three, weighted, dtype):
    from efficientdet_b200 import _lib
    from oracle import graph
    rng = np.random.default_rng(mode * 7 + three)
    B, H, C = 2, 12, 88
    tdt = torch.float32 if dtype == "fp32" else torch.bfloat16
    dt = _lib.F32 if dtype == "fp32" else _lib.BF16
    H0 = {0: H, 1: H // 2, 2: H * 2}[mode]
    a = _dev(rng.standard_normal((B, H0, H0, C)).astype(np.float32), tdt)
    b = _dev(rng.standard_normal((B, H, H, C)).astype(np.float32), tdt)
    c = _dev(rng.standard_normal((B, H, H, C)).astype(np.float32), tdt) if three else None
    fw = np.array([0.7, -0.1, 0.4][:3 if three else 2], np.float32)
    dw = (rng.standard_normal((3, 3, C, 1)) / 3).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, C).astype(np.float32)
    sh = rng.normal(0, 0.2, C).astype(np.float32)
    out = torch.empty((B, H, H, C), dtype=tdt, device="cuda")
    fwd, dwd, scd, shd = _dev(fw), _dev(dw), _dev(sc), _dev(sh)
    _lib.call("effdet_bifpn_node", a.data_ptr(), mode, b.data_ptr(), c.data_ptr() if three else None,
              fwd.data_ptr() if weighted else None, 1e-4, dwd.data_ptr(), scd.data_ptr(), shd.data_ptr(),
              out.data_ptr(), B, H, H, C, dt, _lib.stream_ptr())
    ta = _nchw(a.float().cpu().numpy()).double()
    if mode == 1:
        ta = graph.upsample2(ta)
    elif mode == 2:
        ta = graph.maxpool2(ta)
    ins = [ta, _nchw(b.float().cpu().numpy()).double()] + ([_nchw(c.float().cpu().numpy()).double()] if three else [])
    f = graph.fuse(ins, {"f/f": fw.astype(np.float64)}, weighted, "f")
    y = graph.dwconv2d(f, dw.astype(np.float64), 1)
    y = torch.relu(y * torch.from_numpy(sc).double().view(1, -1, 1, 1) + torch.from_numpy(sh).double().view(1, -1, 1, 1))
    assert rel_err(out.float().cpu().numpy(), _nhwc(y)) < (1e-5 if dtype == "fp32" else 6e-3)


def test_wbifpn_add_layer():
    from efficientdet_b200.layers import wBiFPNAdd
    rng = np.random.default_rng(0)
    xs = [rng.standard_normal((2, 8, 8, 16)).astype(np.float32) for _ in range(3)]
    layer = wBiFPNAdd(name="w_bi_fpn_add_test")
    out = layer(xs)
    w = np.full(3, 1 / 3, np.float32)
    want = (w[0] * xs[0] + w[1] * xs[1] + w[2] * xs[2]) / (w.sum() + np.float32(1e-4))
    assert rel_err(out, want) < 1e-6
    layer.set_weights([np.array([0.5, -1.0, 2.0], np.float32)])
    out = layer(xs)
    want = (0.5 * xs[0] + 2.0 * xs[2]) / (2.5 + 1e-4)
    assert rel_err(out, want) < 1e-6
    assert layer.get_config()["epsilon"] == 1e-4
    assert layer.compute_output_shape([(2, 8, 8, 16)] * 3) == (2, 8, 8, 16)


def test_stem():
    from efficientdet_b200 import _lib
    from oracle import graph
    rng = np.random.default_rng(1)
    B, S, C0 = 2, 34, 32
    x = rng.standard_normal((B, S, S, 3)).astype(np.float32)
    w = (rng.standard_normal((3, 3, 3, C0)) / 5).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, C0).astype(np.float32)
    sh = rng.normal(0, 0.2, C0).astype(np.float32)
    out = torch.empty((B, S // 2, S // 2, C0), dtype=torch.float32, device="cuda")
    xd, wd, scd, shd = _dev(x), _dev(w), _dev(sc), _dev(sh)
    _lib.call("effdet_stem_conv", xd.data_ptr(), wd.data_ptr(), scd.data_ptr(),
              shd.data_ptr(), out.data_ptr(), B, S, S, C0, _lib.F32, _lib.stream_ptr())
    y = graph.conv2d(_nchw(x).double(), w.astype(np.float64), 2)
    y = graph.swish(y * torch.from_numpy(sc).double().view(1, -1, 1, 1) + torch.from_numpy(sh).double().view(1, -1, 1, 1))
    assert rel_err(out.cpu().numpy(), _nhwc(y)) < 1e-5


@pytest.mark.parametrize("C0,S,B,dtype,act", [(32, 34, 2, "fp32", 2), (32, 70, 3, "bf16", 2), (48, 33, 2, "bf16", 2),
                                              (40, 64, 1, "bf16", 0), (16, 20, 2, "fp32", 2), (16, 20, 2, "bf16", 2)])
def test_stem_uint8_input_is_bit_identical(C0, S, B, dtype, act):
    """effdet_stem_conv_u8 (raw letterboxed bytes + the per-byte normalisation table) == the float stem on
    the normalised image (train_tpu.py:135-140), bit for bit; effdet_normalize_u8 == the oracle (pinned on the
    reference's utils.normalize_image by tests/golden/preprocess.npz)."""
    from efficientdet_b200 import _lib
    from efficientdet_b200.utils.preprocess import normalization_lut
    from oracle import preprocess as op
    rng = np.random.default_rng(C0 + S)
    img = rng.integers(0, 256, (B, S, S, 3), dtype=np.uint8)
    norm = op.normalize_image_ref(img)
    w = (rng.standard_normal((3, 3, 3, C0)) / 5).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, C0).astype(np.float32)
    sh = rng.normal(0, 0.2, C0).astype(np.float32)
    Ho = (S + 1) // 2
    tdt = torch.float32 if dtype == "fp32" else torch.bfloat16
    cdt = _lib.F32 if dtype == "fp32" else _lib.BF16
    a = torch.full((B, Ho, Ho, C0), float("nan"), dtype=tdt, device="cuda")
    b = torch.full((B, Ho, Ho, C0), float("nan"), dtype=tdt, device="cuda")
    u8d = torch.from_numpy(img).cuda()
    lut = _dev(normalization_lut())
    xd, wd, scd, shd = _dev(norm), _dev(w), _dev(sc), _dev(sh)
    st = _lib.stream_ptr()
    nf = torch.empty((B, S, S, 3), dtype=torch.float32, device="cuda")
    _lib.call("effdet_normalize_u8", u8d.data_ptr(), lut.data_ptr(), nf.data_ptr(), B * S * S * 3, st)
    assert np.array_equal(nf.cpu().numpy(), norm)
    if dtype == "bf16" and C0 in (32, 40, 48, 56, 64):
        _lib.call("effdet_stem_conv_act", xd.data_ptr(), wd.data_ptr(), scd.data_ptr(), shd.data_ptr(),
                  a.data_ptr(), B, S, S, C0, act, st)
    else:
        _lib.call("effdet_stem_conv", xd.data_ptr(), wd.data_ptr(), scd.data_ptr(), shd.data_ptr(),
                  a.data_ptr(), B, S, S, C0, cdt, st)
    _lib.call("effdet_stem_conv_u8", u8d.data_ptr(), lut.data_ptr(), wd.data_ptr(), scd.data_ptr(),
              shd.data_ptr(), b.data_ptr(), B, S, S, C0, act, cdt, st)
    torch.cuda.synchronize()
    assert np.array_equal(a.view(torch.int16 if dtype == "bf16" else torch.int32).cpu().numpy(),
                          b.view(torch.int16 if dtype == "bf16" else torch.int32).cpu().numpy())
    assert not torch.isnan(b.float()).any()


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_model_accepts_raw_uint8_images(dtype):
    """predict_on_batch(uint8 letterboxed image) == predict_on_batch(normalize_image(image)), bit-exact, through
    FilterDetections; the uint8 image comes from utils.preprocess.preprocess_image (utils/__init__.py:103-138)."""
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.utils.anchors import anchors_for_shape
    from efficientdet_b200.utils.preprocess import normalize_image, preprocess_image
    size, classes = 128, 5
    anchors = anchors_for_shape((size, size))
    model, pmodel = efficientdet(0, num_classes=classes, image_size=size, score_threshold=0.3, dtype=dtype,
                                 drop_connect_rate=0, anchors=anchors)
    perturb_weights(model)
    rng = np.random.default_rng(3)
    raw = [rng.integers(0, 256, (97, 150, 3), dtype=np.uint8), rng.integers(0, 256, (200, 131, 3), dtype=np.uint8)]
    boxed = np.stack([preprocess_image(r, size)[0] for r in raw])
    assert boxed.dtype == np.uint8 and boxed.shape == (2, size, size, 3)
    r8, c8 = model.predict_on_batch(boxed)
    rf, cf = model.predict_on_batch(normalize_image(boxed))
    assert np.array_equal(r8, rf) and np.array_equal(c8, cf)
    d8 = pmodel.predict_on_batch([boxed])
    df = pmodel.predict_on_batch([normalize_image(boxed)])
    for x, y in zip(d8, df):
        assert np.array_equal(x, y)
    g8 = list(pmodel.predict_generator([torch.from_numpy(boxed).pin_memory()]))[0]
    for x, y in zip(g8, df):
        assert np.array_equal(x, y)


# ------------------------------------------------------------------ whole network
@pytest.mark.parametrize("phi,size,weighted,dtype,classes", [
    (0, 128, False, "fp32", 20), (0, 256, True, "fp32", 90), (0, 256, True, "bf16", 20),
    (1, 128, True, "fp32", 8), (2, 128, False, "bf16", 8)])
def test_network_forward_per_level(phi, size, weighted, dtype, classes):
    from efficientdet_b200.model import efficientdet
    from oracle import graph
    model = efficientdet(phi, num_classes=classes, weighted_bifpn=weighted, image_size=size,
                         dtype=dtype, drop_connect_rate=0, just_training_model=True)
    W = perturb_weights(model)
    rng = np.random.default_rng(1234)
    B = 2
    img = rng.standard_normal((B, size, size, 3)).astype(np.float32)
    plan = model.net.plan(B, keep_taps=True)
    reg, cls = plan.forward(torch.from_numpy(img).cuda())
    torch.cuda.synchronize()
    taps = {}
    with torch.no_grad():
        r0, c0 = graph.forward(W, img, phi, classes, weighted, taps=taps)
    tol = FP32_TOL if dtype == "fp32" else BF16_TOL
    worst = {}
    for name in ["C3", "C4", "C5"] + ["BiFPN_%d_P%d" % (i, l) for i in range(2 + phi) for l in range(3, 8)]:
        got = plan.tensor(plan.taps[name]).float().cpu().numpy()
        worst[name] = rel_err(got, taps[name].numpy())
    worst["regression"] = rel_err(reg.cpu().numpy(), r0.numpy())
    worst["classification"] = rel_err(cls.cpu().numpy(), c0.numpy())
    bad = {k: v for k, v in worst.items() if not v < tol}
    assert not bad, (bad, worst)
    # model-level API (buffer-reusing plan) returns bit-identical numbers: the forward pass
    # is deterministic (no float atomics)
    r1, c1 = model.predict_on_batch(img)
    assert np.array_equal(r1, reg.cpu().numpy()) and np.array_equal(c1, cls.cpu().numpy())


def test_prediction_model_end_to_end():
    """efficientdet() -> prediction_model.predict_on_batch([images, anchors]) vs the oracle tail
    applied to the oracle's own regression/classification is checked stage-wise: the GPU tail on
    the GPU's head outputs must be bit-exact w.r.t. the oracle tail on the same inputs."""
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.utils.anchors import anchors_for_shape
    from oracle import tail
    size, classes = 256, 6
    anchors = anchors_for_shape((size, size))
    model, pmodel = efficientdet(0, num_classes=classes, image_size=size, score_threshold=0.4,
                                 drop_connect_rate=0)
    perturb_weights(model)
    _, pbaked = efficientdet(0, num_classes=classes, image_size=size, score_threshold=0.4,
                             drop_connect_rate=0, anchors=anchors)
    pbaked.set_weights_dict({k: v for k, v in model.get_weights_dict().items()})
    rng = np.random.default_rng(7)
    img = rng.standard_normal((2, size, size, 3)).astype(np.float32)
    reg, cls = model.predict_on_batch(img)
    boxes, scores, labels = pmodel.predict_on_batch([img, anchors[None].astype(np.float32)])
    b2, s2, l2 = pbaked.predict_on_batch([img])
    wb = tail.clip_boxes((2, size, size, 3), tail.apply_bbox_deltas(anchors[None].astype(np.float32), reg))
    ob, os_, ol = tail.filter_detections_batch(wb, cls, score_threshold=0.4)
    assert np.array_equal(boxes, ob) and np.array_equal(scores, os_) and np.array_equal(labels, ol)
    assert np.array_equal(b2, ob) and np.array_equal(s2, os_) and np.array_equal(l2, ol)
    assert boxes.shape == (2, 300, 4) and labels.dtype == np.int32


def test_predict_generator_matches_predict_on_batch():
    """The prefetching pipeline (host->device copy of batch i+1 on a copy stream while batch i runs)
    returns, batch by batch, exactly what predict_on_batch returns."""
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.utils.anchors import anchors_for_shape
    size, classes = 128, 4
    anchors = anchors_for_shape((size, size))
    model, pmodel = efficientdet(0, num_classes=classes, image_size=size, score_threshold=0.3,
                                 drop_connect_rate=0, anchors=anchors, dtype="bf16")
    perturb_weights(model)
    rng = np.random.default_rng(3)
    batches = [rng.standard_normal((3, size, size, 3)).astype(np.float32) for _ in range(5)]
    want = [pmodel.predict_on_batch([b]) for b in batches]
    got = list(pmodel.predict_generator(torch.from_numpy(b).pin_memory() for b in batches))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert np.array_equal(a, b)


def test_keras_h5_weight_file_round_trip_through_the_model(tmp_path):
    """model.save_weights('x.h5') -> Keras weight-file layout (utils/hdf5.py) -> load_weights(by_name=True) on a
    freshly initialised model (train.py:329-332): identical weights and identical detections."""
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.utils import hdf5
    from efficientdet_b200.utils.anchors import anchors_for_shape
    size = 128
    anchors = anchors_for_shape((size, size))
    m1, p1 = efficientdet(0, num_classes=4, weighted_bifpn=True, image_size=size, drop_connect_rate=0,
                          score_threshold=0.3, anchors=anchors, seed=1)
    perturb_weights(m1)
    path = str(tmp_path / "weights.h5")
    m1.save_weights(path)
    root = hdf5.open_file(path)
    layers = [x.decode() for x in root.attrs["layer_names"]]
    assert "stem_conv" in layers and "box_head" in layers and "w_bi_fpn_add" in layers
    assert [x.decode() for x in root["stem_bn"].attrs["weight_names"]] == [
        "stem_bn/gamma:0", "stem_bn/beta:0", "stem_bn/moving_mean:0", "stem_bn/moving_variance:0"]
    m2, p2 = efficientdet(0, num_classes=4, weighted_bifpn=True, image_size=size, drop_connect_rate=0,
                          score_threshold=0.3, anchors=anchors, seed=2)
    m2.load_weights(path, by_name=True)
    w1, w2 = m1.get_weights_dict(), m2.get_weights_dict()
    assert set(w1) == set(w2) and all(np.array_equal(w1[k], w2[k]) for k in w1)
    img = np.random.default_rng(0).standard_normal((2, size, size, 3)).astype(np.float32)
    for a, b in zip(p1.predict_on_batch([img]), p2.predict_on_batch([img])):
        assert np.array_equal(a, b)


def test_letterbox_resize_on_device_bit_exact():
    """effdet_letterbox_u8 (utils.preprocess.preprocess_images_device) == the reference's utils.resize_image
    (cv2.resize bilinear + grey canvas): the golden vectors produced by executing the reference, then random image
    sizes at the model's input sizes against the oracle restatement (itself pinned on the same vectors)."""
    import os
    from efficientdet_b200.utils.preprocess import preprocess_images_device
    from oracle import preprocess as op
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess.npz"))
    for i, (h, w, size) in enumerate(gold["cases"]):
        out, meta = preprocess_images_device([gold["img_%d" % i]], int(size))
        assert np.array_equal(out[0].cpu().numpy(), gold["boxed_%d" % i]), (i, h, w, size)
        assert np.array_equal(np.array(meta[0], np.float64), gold["meta_%d" % i])
    rng = np.random.default_rng(99)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            for h, w in [(480, 640), (375, 500), (1080, 1920), (333, 500), (512, 512), (640, 427), (97, 1400), (512, 300)]]
    for size in (512, 768):
        out, meta = preprocess_images_device(imgs, size)
        for j, img in enumerate(imgs):
            want, scale, oh, ow = op.resize_image_ref(img, size)
            assert np.array_equal(out[j].cpu().numpy(), want), (size, img.shape)
            assert meta[j] == (scale, oh, ow)


@pytest.mark.parametrize("k,stride,C,H", [(3, 1, 48, 20), (5, 1, 96, 17), (3, 2, 144, 22), (5, 2, 40, 19)])
def test_dwconv_split_output_equals_dwconv_then_split(k, stride, C, H):
    """fp32 accuracy mode on the tensor cores: the depthwise kernel writing the bf16 hi | lo planes directly
    (effdet_dwconv_split_out) == effdet_dwconv followed by effdet_split_bf16, bit for bit, SE partial sums included."""
    from efficientdet_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(k * 100 + C)
    B = 2
    Ho = -(-H // stride)
    x = torch.from_numpy(rng.standard_normal((B, H, H, C)).astype(np.float32)).cuda()
    w = torch.from_numpy((rng.standard_normal((k, k, C)) * 0.3).astype(np.float32)).cuda()
    sc = torch.from_numpy(rng.uniform(0.5, 1.5, C).astype(np.float32)).cuda()
    sh = torch.from_numpy(rng.normal(0, 0.2, C).astype(np.float32)).cuda()
    nblk = lib.effdet_dwconv_se_blocks(B, H, H, C, stride, _lib.F32)
    y = torch.empty((B, Ho, Ho, C), device="cuda")
    p1 = torch.empty((B, nblk, C), device="cuda")
    p2 = torch.empty((B, nblk, C), device="cuda")
    st = _lib.stream_ptr()
    _lib.call("effdet_dwconv", x.data_ptr(), w.data_ptr(), sc.data_ptr(), sh.data_ptr(), y.data_ptr(), p1.data_ptr(),
              nblk, B, H, H, C, k, stride, _lib.ACT_SWISH, _lib.F32, st)
    want = torch.empty((B, Ho, Ho, 2 * C), device="cuda", dtype=torch.bfloat16)
    _lib.call("effdet_split_bf16", y.data_ptr(), want.data_ptr(), B * Ho * Ho, C, st)
    got = torch.full((B, Ho, Ho, 2 * C), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.call("effdet_dwconv_split_out", x.data_ptr(), w.data_ptr(), sc.data_ptr(), sh.data_ptr(), got.data_ptr(),
              p2.data_ptr(), nblk, B, H, H, C, k, stride, _lib.ACT_SWISH, st)
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    assert torch.equal(p1, p2)
