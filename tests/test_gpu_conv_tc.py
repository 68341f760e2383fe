"""GPU parity of the tcgen05/TMA implicit-GEMM convolution (bf16 operands, fp32 accumulate in
TMEM) against the fp64 oracle convolution evaluated on the SAME bf16-rounded inputs/weights."""
import ctypes

import numpy as np
import pytest
import torch

from util_model import rel_err

pytestmark = pytest.mark.gpu


def _d(a, dt=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)


def _bf(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy()


def _panel(w, mode, gate=None, B=0):
    from efficientdet_b200 import _lib
    lib = _lib.load()
    k = w.shape[0]
    taps, cin, cout = k * k, w.shape[2], w.shape[3]
    K, N = (cin, cout) if mode == 0 else (cout, cin)
    n = lib.effdet_conv_weight_panel_elems(B if gate is not None else taps, K, N)
    panel = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    wd = _d(w)
    gd = _d(gate) if gate is not None else None
    _lib.call("effdet_conv_weight_panel", wd.data_ptr(), panel.data_ptr(), taps, cin, cout, mode,
              gd.data_ptr() if gate is not None else None, B, _lib.stream_ptr())
    torch.cuda.synchronize()
    return panel


def _run(xs, panel, cin, cout, k, B, out_dtype, act=0, scale=None, shift=None, res=None, mask=None,
         keep=None, per_sample=False, ys=None, ldc=None, ybs=None, gate=None):
    from efficientdet_b200 import _lib
    d = _lib.ConvDesc()
    d.n_groups = len(xs)
    outs = []
    for i, x in enumerate(xs):
        H = x.shape[1]
        if ys is None:
            y = torch.full((B, H, H, cout), float("nan"), device="cuda",
                           dtype=torch.float32 if out_dtype == _lib.F32 else torch.bfloat16)
        else:
            y = ys[i]
        outs.append(y)
        d.x[i], d.y[i] = x.data_ptr(), y.data_ptr()
        d.residual[i] = res[i].data_ptr() if res else None
        d.relu_mask[i] = mask[i].data_ptr() if mask else None
        d.H[i] = d.W[i] = H
        if ldc:
            d.ldc[i], d.y_batch_stride[i] = ldc[i], ybs[i]
    d.B, d.Cin, d.Cout, d.kh, d.kw, d.stride = B, cin, cout, k, k, 1
    d.weight = None
    d.scale = scale.data_ptr() if scale is not None else None
    d.shift = shift.data_ptr() if shift is not None else None
    d.keep = keep.data_ptr() if keep is not None else None
    d.gate = gate.data_ptr() if gate is not None else None
    d.act, d.in_dtype, d.out_dtype = act, _lib.BF16, out_dtype
    d.weight_bf16, d.allow_tensor_core, d.weight_per_sample = panel.data_ptr(), 1, int(per_sample)
    _lib.call("effdet_conv2d", ctypes.byref(d), _lib.stream_ptr())
    torch.cuda.synchronize()
    return outs


def _ref(x_bf, w, act=0, scale=None, shift=None):
    from oracle import graph
    y = graph.conv2d(torch.from_numpy(x_bf).double().permute(0, 3, 1, 2), _bf(w).astype(np.float64), 1)
    if scale is not None:
        y = y * torch.from_numpy(scale).double().view(1, -1, 1, 1)
    if shift is not None:
        y = y + torch.from_numpy(shift).double().view(1, -1, 1, 1)
    y = [lambda v: v, torch.relu, graph.swish, torch.sigmoid][act](y)
    return y.permute(0, 2, 3, 1).numpy()


@pytest.mark.parametrize("cin,cout,H,B,act", [(64, 64, 16, 2, 1), (24, 144, 16, 3, 2), (320, 64, 8, 2, 1),
                                              (1152, 320, 4, 2, 0), (40, 240, 20, 1, 2), (88, 88, 10, 5, 1),
                                              # D4 / D6 backbone widths (block 7: 448 -> 2688 -> 448, 576 -> 3456 -> 576)
                                              (2688, 448, 8, 2, 0), (448, 2688, 8, 1, 2), (3456, 576, 6, 1, 0),
                                              (576, 3456, 5, 2, 2), (1632, 272, 16, 2, 0)])
def test_conv1x1_tc(cin, cout, H, B, act):
    from efficientdet_b200 import _lib
    rng = np.random.default_rng(cin + cout)
    x = rng.standard_normal((B, H, H, cin)).astype(np.float32)
    w = (rng.standard_normal((1, 1, cin, cout)) / np.sqrt(cin)).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, cout).astype(np.float32); sh = rng.normal(0, 0.2, cout).astype(np.float32)
    xd = _d(x, torch.bfloat16)
    res = rng.standard_normal((B, H, H, cout)).astype(np.float32)
    keep = rng.uniform(0.5, 1.5, B).astype(np.float32)
    resd, keepd = _d(res, torch.bfloat16), _d(keep)
    y, = _run([xd], _panel(w, 0), cin, cout, 1, B, _lib.BF16, act, _d(sc), _d(sh), res=[resd], keep=keepd)
    want = _ref(xd.float().cpu().numpy(), w, act, sc, sh) * keep[:, None, None, None] + resd.float().cpu().numpy()
    assert rel_err(y.float().cpu().numpy(), want) < 6e-3


@pytest.mark.parametrize("cin,cout,H,B,k", [(320, 64, 16, 3, 3), (64, 64, 8, 2, 3), (64, 64, 9, 2, 3), (40, 48, 14, 2, 1)])
def test_conv_stride2_tc(cin, cout, H, B, k):
    """BiFPN P6/P7 laterals (model.py:205-211): stride 2, TF SAME padding (asymmetric on even inputs),
    sampled through the input tensor map's element strides."""
    from efficientdet_b200 import _lib
    from oracle import graph
    rng = np.random.default_rng(cin + H)
    x = rng.standard_normal((B, H, H, cin)).astype(np.float32)
    w = (rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, cout).astype(np.float32); sh = rng.normal(0, 0.2, cout).astype(np.float32)
    xd = _d(x, torch.bfloat16)
    Ho = (H + 1) // 2
    y = torch.full((B, Ho, Ho, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    d = _lib.ConvDesc()
    d.n_groups = 1
    d.x[0], d.y[0] = xd.data_ptr(), y.data_ptr()
    d.H[0] = d.W[0] = H
    d.B, d.Cin, d.Cout, d.kh, d.kw, d.stride = B, cin, cout, k, k, 2
    scd, shd = _d(sc), _d(sh)
    d.scale, d.shift = scd.data_ptr(), shd.data_ptr()
    d.act, d.in_dtype, d.out_dtype = 1, _lib.BF16, _lib.BF16
    panel = _panel(w, 0)
    d.weight_bf16, d.allow_tensor_core = panel.data_ptr(), 1
    n0 = _lib.launch_count()
    _lib.call("effdet_conv2d", ctypes.byref(d), _lib.stream_ptr())
    torch.cuda.synchronize()
    assert _lib.launch_count() == n0 + 1
    ref = graph.conv2d(xd.float().cpu().double().permute(0, 3, 1, 2), _bf(w).astype(np.float64), 2)
    ref = torch.relu(ref * torch.from_numpy(sc).double().view(1, -1, 1, 1) +
                     torch.from_numpy(sh).double().view(1, -1, 1, 1)).permute(0, 2, 3, 1).numpy()
    assert rel_err(y.float().cpu().numpy(), ref) < 6e-3


def test_conv1x1_tc_per_sample_gate():
    from efficientdet_b200 import _lib
    rng = np.random.default_rng(9)
    B, H, cin, cout = 3, 16, 144, 40
    x = rng.standard_normal((B, H, H, cin)).astype(np.float32)
    w = (rng.standard_normal((1, 1, cin, cout)) / np.sqrt(cin)).astype(np.float32)
    gate = rng.uniform(0.1, 1.0, (B, cin)).astype(np.float32)
    xd, gd = _d(x, torch.bfloat16), _d(gate)
    y, = _run([xd], _panel(w, 0, gate, B), cin, cout, 1, B, _lib.BF16, per_sample=True, gate=gd)
    xq = xd.float().cpu().numpy().astype(np.float64)
    wq = np.stack([_bf(w[0, 0] * gate[b][:, None]) for b in range(B)]).astype(np.float64)
    want = np.einsum("bhwc,bcn->bhwn", xq, wq)
    assert rel_err(y.float().cpu().numpy(), want) < 6e-3


@pytest.mark.parametrize("W,cout,out_f32,B,sizes", [(64, 64, False, 4, [16, 8, 4, 2, 1]),
                                                    (64, 36, True, 2, [16, 8, 4, 2, 1]),
                                                    # 6 / 7 classes on a 64-wide BiFPN: fp32 outputs of 49..64
                                                    # channels exceed the halo form's shared memory (ring form)
                                                    (64, 54, True, 2, [16, 8, 4, 2, 1]),
                                                    (64, 63, True, 2, [32, 16, 8, 4, 2]),
                                                    (64, 48, True, 2, [16, 8, 4, 2, 1]),
                                                    (88, 180, True, 2, [20, 10, 5]),
                                                    (112, 810, True, 1, [12, 6, 3]),
                                                    # D4 (W 224, head depth 4) and D6 / D7 (W 384) head widths
                                                    (224, 224, False, 2, [32, 16, 8, 4, 2]),
                                                    (224, 810, True, 1, [16, 8, 4, 2, 1]),
                                                    (384, 384, False, 1, [22, 11, 6, 3, 2]),
                                                    (384, 36, True, 2, [11, 6, 3])])
def test_conv3x3_head_tc_grouped(W, cout, out_f32, B, sizes):
    """All pyramid levels in one launch; fp32 outputs land in the concatenated (B, N, per) layout."""
    from efficientdet_b200 import _lib
    rng = np.random.default_rng(W + cout)
    w = (rng.standard_normal((3, 3, W, cout)) / np.sqrt(9 * W)).astype(np.float32)
    bias = rng.normal(0, 0.2, cout).astype(np.float32)
    xs = [_d(rng.standard_normal((B, s, s, W)).astype(np.float32), torch.bfloat16) for s in sizes]
    act = 3 if out_f32 and cout != 36 else (0 if out_f32 else 1)
    if out_f32:
        rows = sum(s * s for s in sizes)
        out = torch.full((B, rows, cout), float("nan"), device="cuda")
        offs = np.concatenate([[0], np.cumsum([s * s for s in sizes])[:-1]])
        ys = [out.view(-1)[int(o) * cout:] for o in offs]
        _run(xs, _panel(w, 0), W, cout, 3, B, _lib.F32, act, None, _d(bias), ys=ys, ldc=[cout] * len(sizes),
             ybs=[rows * cout] * len(sizes))
        got = out.cpu().numpy()
        want = np.concatenate([_ref(x.float().cpu().numpy(), w, act, None, bias).reshape(B, -1, cout) for x in xs], 1)
        assert rel_err(got, want) < 6e-3
    else:
        ys = _run(xs, _panel(w, 0), W, cout, 3, B, _lib.BF16, act, None, _d(bias))
        for x, y in zip(xs, ys):
            assert rel_err(y.float().cpu().numpy(), _ref(x.float().cpu().numpy(), w, act, None, bias)) < 6e-3


def test_conv3x3_dgrad_tc_with_relu_mask_and_accumulate():
    from efficientdet_b200 import _lib
    from oracle import graph
    rng = np.random.default_rng(3)
    B, H, cin, cout = 2, 12, 64, 64
    w = (rng.standard_normal((3, 3, cin, cout)) / np.sqrt(9 * cin)).astype(np.float32)
    dz = _d(rng.standard_normal((B, H, H, cout)).astype(np.float32), torch.bfloat16)
    mask = _d(rng.standard_normal((B, H, H, cin)).astype(np.float32), torch.bfloat16)
    prev = _d(rng.standard_normal((B, H, H, cin)).astype(np.float32), torch.bfloat16)
    x = torch.zeros((B, cin, H, H), dtype=torch.float64, requires_grad=True)
    graph.conv2d(x, _bf(w).astype(np.float64), 1).backward(dz.double().cpu().permute(0, 3, 1, 2))
    want = x.grad.permute(0, 2, 3, 1).numpy() * (mask.float().cpu().numpy() > 0) + prev.float().cpu().numpy()
    out = prev.clone()
    _run([dz], _panel(w, 1), cout, cin, 3, B, _lib.BF16, res=[out], mask=[mask], ys=[out])
    assert rel_err(out.float().cpu().numpy(), want) < 6e-3


@pytest.mark.parametrize("k,cin,cout,ldz,sizes,B", [(3, 64, 64, 64, [16, 8, 4, 2, 1], 4), (1, 112, 64, 64, [16], 2),
                                                      (3, 64, 36, 40, [8, 4], 3), (1, 320, 64, 64, [8], 2),
                                                      (3, 88, 180, 184, [10, 5], 2), (3, 64, 810, 816, [4], 1),
                                                      # D4 / D6 head widths: the column-of-taps form (Cin > 64)
                                                      (3, 224, 224, 224, [32, 16, 8, 4, 2], 2),
                                                      (3, 224, 810, 816, [16, 8], 1), (3, 224, 36, 40, [16, 4], 3),
                                                      (3, 384, 384, 384, [22, 11], 1), (3, 160, 160, 160, [14, 7], 2)])
def test_conv_wgrad_tc(k, cin, cout, ldz, sizes, B):
    """tcgen05 weight gradient with MN-major TMA-fed operands vs fp64 autograd on the same
    bf16-rounded operands (dz lives in channel-padded per-level buffers)."""
    from efficientdet_b200 import _lib
    from oracle import graph
    lib = _lib.load()
    rng = np.random.default_rng(k * 1000 + cin + cout)
    w = torch.zeros((k, k, cin, cout), dtype=torch.float64, requires_grad=True)
    d = _lib.WgradDesc()
    d.n_groups = len(sizes)
    keep, total = [], 0
    for i, H in enumerate(sizes):
        x = _d(rng.standard_normal((B, H, H, cin)).astype(np.float32), torch.bfloat16)
        dz = torch.zeros((B, H, H, ldz), dtype=torch.bfloat16, device="cuda")
        dz[..., :cout] = _d(rng.standard_normal((B, H, H, cout)).astype(np.float32), torch.bfloat16)
        dz[..., cout:] = 7.0          # padding channels must be ignored
        keep += [x, dz]
        y = graph.conv2d(x.double().cpu().permute(0, 3, 1, 2), w, 1)
        total = total + (y * dz[..., :cout].double().cpu().permute(0, 3, 1, 2)).sum()
        d.x[i], d.dz[i], d.H[i], d.W[i], d.dz_ld[i] = x.data_ptr(), dz.data_ptr(), H, H, ldz
    total.backward()
    d.B, d.Cin, d.Cout, d.kh, d.kw, d.stride = B, cin, cout, k, k, 1
    d.x_dtype = d.dz_dtype = _lib.BF16
    ns = lib.effdet_conv_wgrad_tc_splits(ctypes.byref(d))
    assert ns > 0
    fuse_bias = bool(lib.effdet_conv_wgrad_tc_fuses_bias(ctypes.byref(d)))
    assert fuse_bias == (k == 3 and cin <= 64)
    part = torch.full((ns * (k * k * cin * cout + (cout if fuse_bias else 0)),), float("nan"), device="cuda")
    out = torch.full((k, k, cin, cout), float("nan"), device="cuda")
    dbias = torch.full((cout,), float("nan"), device="cuda")
    d.dweight, d.partial, d.n_splits, d.accumulate = out.data_ptr(), part.data_ptr(), ns, 0
    d.dbias = dbias.data_ptr() if fuse_bias else None
    _lib.call("effdet_conv_wgrad_tc", ctypes.byref(d), _lib.stream_ptr())
    torch.cuda.synchronize()
    assert rel_err(out.cpu().numpy(), w.grad.numpy()) < 1e-4
    if fuse_bias:
        # bias gradient from the same launch (ones block in the spare half of the last tap pair)
        want = sum(keep[2 * i + 1][..., :cout].double().sum((0, 1, 2)) for i in range(len(sizes))).cpu().numpy()
        assert rel_err(dbias.cpu().numpy(), want) < 1e-4


# ------------------------------------------------------------------ fp32 accuracy mode on the tensor cores
def _split(x):
    """fp32 (B,H,W,C) cuda -> bf16 (B,H,W,2C) hi | lo planes (effdet_split_bf16)."""
    from efficientdet_b200 import _lib
    B, H, W, C = x.shape
    out = torch.empty((B, H, W, 2 * C), dtype=torch.bfloat16, device="cuda")
    _lib.call("effdet_split_bf16", x.data_ptr(), out.data_ptr(), B * H * W, C, _lib.stream_ptr())
    torch.cuda.synchronize()
    return out


def _panel_split(w, gate=None, B=0):
    from efficientdet_b200 import _lib
    lib = _lib.load()
    k = w.shape[0]
    taps, cin, cout = k * k, w.shape[2], w.shape[3]
    n = lib.effdet_conv_weight_panel_split_elems(B if gate is not None else taps, cin, cout)
    panel = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    wd = _d(w)
    gd = _d(gate) if gate is not None else None
    _lib.call("effdet_conv_weight_panel_split", wd.data_ptr(), panel.data_ptr(), taps, cin, cout,
              gd.data_ptr() if gate is not None else None, B, _lib.stream_ptr())
    torch.cuda.synchronize()
    return panel


def _run_split(xs, panel, cin, cout, k, B, stride=1, act=0, scale=None, shift=None, res=None, per_sample=False):
    from efficientdet_b200 import _lib
    d = _lib.ConvDesc()
    d.n_groups = len(xs)
    outs = []
    for i, x in enumerate(xs):
        H = x.shape[1]
        Ho = -(-H // stride)
        y = torch.full((B, Ho, Ho, cout), float("nan"), device="cuda", dtype=torch.float32)
        outs.append(y)
        d.x[i], d.y[i] = x.data_ptr(), y.data_ptr()
        d.residual[i] = res[i].data_ptr() if res else None
        d.H[i] = d.W[i] = H
    d.B, d.Cin, d.Cout, d.kh, d.kw, d.stride = B, cin, cout, k, k, stride
    d.scale = scale.data_ptr() if scale is not None else None
    d.shift = shift.data_ptr() if shift is not None else None
    d.act, d.in_dtype, d.out_dtype = act, _lib.BF16, _lib.F32
    d.weight_bf16, d.allow_tensor_core, d.weight_per_sample, d.split_planes = panel.data_ptr(), 1, int(per_sample), 1
    _lib.call("effdet_conv2d", ctypes.byref(d), _lib.stream_ptr())
    torch.cuda.synchronize()
    return outs


def test_split_bf16_planes():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((2, 5, 7, 24)) * np.exp(rng.normal(0, 3, (2, 5, 7, 24)))).astype(np.float32)
    s = _split(_d(x)).float().cpu().numpy()
    hi, lo = s[..., :24], s[..., 24:]
    assert np.array_equal(hi, _bf(x))
    assert np.abs(x.astype(np.float64) - (hi.astype(np.float64) + lo)).max() <= 2.0 ** -17 * np.abs(x).max()
    assert (np.abs(x - (hi + lo)) <= 2.0 ** -16 * np.abs(x) + 1e-38).all()


@pytest.mark.parametrize("cin,cout,H,B,k,stride,act", [
    (24, 144, 16, 3, 1, 1, 2), (1152, 320, 4, 2, 1, 1, 0), (40, 240, 20, 1, 1, 1, 2), (64, 64, 16, 2, 3, 1, 1),
    (112, 112, 12, 2, 3, 1, 1), (320, 64, 16, 3, 3, 2, 0), (88, 36, 10, 2, 3, 1, 0)])
def test_conv_fp32_on_tensor_cores(cin, cout, H, B, k, stride, act):
    """split_planes: fp32 activations and weights through the bf16 tensor-core kernel as the three-term product
    hi*Whi + lo*Whi + hi*Wlo, fp32 accumulate -- against the fp64 convolution of the UNROUNDED fp32 operands.
    Bound 2e-5 (max-normalised; each product is good to ~2^-18, the epilogue's swish to ~2^-21)."""
    rng = np.random.default_rng(cin * 7 + cout + k)
    x = rng.standard_normal((B, H, H, cin)).astype(np.float32)
    w = (rng.standard_normal((k, k, cin, cout)) / np.sqrt(cin * k * k)).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, cout).astype(np.float32); sh = rng.normal(0, 0.2, cout).astype(np.float32)
    Ho = -(-H // stride)
    res = rng.standard_normal((B, Ho, Ho, cout)).astype(np.float32)
    y, = _run_split([_split(_d(x))], _panel_split(w), cin, cout, k, B, stride, act, _d(sc), _d(sh), res=[_d(res)])
    from oracle import graph
    r = graph.conv2d(torch.from_numpy(x).double().permute(0, 3, 1, 2), w.astype(np.float64), stride)
    r = r * torch.from_numpy(sc).double().view(1, -1, 1, 1) + torch.from_numpy(sh).double().view(1, -1, 1, 1)
    r = [lambda v: v, torch.relu, graph.swish, torch.sigmoid][act](r).permute(0, 2, 3, 1).numpy() + res
    assert rel_err(y.cpu().numpy(), r) < 2e-5


def test_conv_fp32_on_tensor_cores_gated_and_grouped():
    """per-sample split panels (squeeze-excite gate folded in) and five pyramid levels in one launch."""
    rng = np.random.default_rng(5)
    B, cin, cout = 3, 144, 24
    x = rng.standard_normal((B, 12, 12, cin)).astype(np.float32)
    w = (rng.standard_normal((1, 1, cin, cout)) / np.sqrt(cin)).astype(np.float32)
    gate = rng.uniform(0, 1, (B, cin)).astype(np.float32)
    y, = _run_split([_split(_d(x))], _panel_split(w, gate, B), cin, cout, 1, B, per_sample=True)
    want = np.einsum("bhwc,bc,co->bhwo", x.astype(np.float64), gate.astype(np.float64), w[0, 0].astype(np.float64))
    assert rel_err(y.cpu().numpy(), want) < 2e-5
    from oracle import graph
    w3 = (rng.standard_normal((3, 3, 64, 64)) / 24).astype(np.float32)
    xs = [rng.standard_normal((2, h, h, 64)).astype(np.float32) for h in (16, 8, 4, 2, 1)]
    ys = _run_split([_split(_d(v)) for v in xs], _panel_split(w3), 64, 64, 3, 2, act=1)
    for v, y in zip(xs, ys):
        r = torch.relu(graph.conv2d(torch.from_numpy(v).double().permute(0, 3, 1, 2), w3.astype(np.float64), 1))
        assert rel_err(y.cpu().numpy(), r.permute(0, 2, 3, 1).numpy()) < 2e-5
