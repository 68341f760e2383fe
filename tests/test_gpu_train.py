"""GPU parity of the training step (losses, backward through heads + BiFPN, BN batch statistics,
SGD) against the torch-CPU autograd oracle.  fp32 mode; gradients compared per tensor with
max|got-want| / max|want| <= 2e-3 (the forward tolerance of BASELINE.json is 1e-4; backward
sums are ~30x longer)."""
import numpy as np
import pytest
import torch

from util_model import perturb_weights, rel_err

pytestmark = pytest.mark.gpu


def _targets(size, B, C, seed=7):
    from oracle import anchors as oa
    rng = np.random.default_rng(seed)
    anchors = oa.anchors_for_shape((size, size))
    ann = []
    for _ in range(B):
        n = int(rng.integers(1, 6))
        wh = rng.uniform(size * 0.1, size * 0.5, (n, 2))
        xy = rng.uniform(0, size * 0.5, (n, 2))
        ann.append({"bboxes": np.concatenate([xy, xy + wh], 1).astype(np.float32),
                    "labels": rng.integers(0, C, n).astype(np.float32)})
    reg_t, lab_t = oa.anchor_targets_bbox(anchors, [(size, size, 3)] * B, ann, C)
    return anchors, ann, reg_t, lab_t


def test_losses_fwd_bwd():
    from efficientdet_b200 import _lib
    from oracle import losses
    B, N, C = 2, 3000, 7
    rng = np.random.default_rng(0)
    p = rng.uniform(0.001, 0.999, (B, N, C)).astype(np.float32)
    p[0, :5, 0] = [0.0, 1.0, 1e-9, 1 - 1e-9, 0.5]
    reg = rng.normal(0, 1.2, (B, N, 4)).astype(np.float32)
    state = rng.choice([-1, 0, 1], (B, N), p=[0.1, 0.8, 0.1]).astype(np.float32)
    clsid = rng.integers(0, C, (B, N))
    lab = np.zeros((B, N, C + 1), np.float32)
    bi, ni = np.nonzero(state == 1)
    lab[bi, ni, clsid[bi, ni]] = 1
    lab[..., C] = state
    reg_t = np.concatenate([rng.normal(0, 1, (B, N, 4)), state[..., None]], -1).astype(np.float32)
    pt = torch.tensor(p, dtype=torch.float64, requires_grad=True)
    rt = torch.tensor(reg, dtype=torch.float64, requires_grad=True)
    fl = losses.focal(torch.tensor(lab, dtype=torch.float64), pt, 0.25, 1.5)
    sl = losses.smooth_l1(torch.tensor(reg_t, dtype=torch.float64), rt)
    (fl + sl).backward()
    want_dlogit = (pt.grad * pt.detach() * (1 - pt.detach())).numpy()
    lib = _lib.load()
    d = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)
    pd, rd, rtd, labd = d(p), d(reg), d(reg_t), d(lab)
    dcls = torch.empty((B, N, C), device="cuda"); dreg = torch.empty((B, N, 4), device="cuda")
    out8 = torch.zeros(8, device="cuda")
    wsb = lib.effdet_detection_losses_workspace_size()
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    for dense in (True, False):
        st = d(state, torch.int8); cl = d(np.where(state == 1, clsid, -1), torch.int32)
        _lib.call("effdet_detection_losses", pd.data_ptr(), rd.data_ptr(), rtd.data_ptr(),
                  labd.data_ptr() if dense else None, st.data_ptr(), cl.data_ptr(), B, N, C, 0.25, 1.5,
                  1.0, 1.0, dcls.data_ptr(), dreg.data_ptr(), out8.data_ptr(), ws.data_ptr(), wsb,
                  None, None, None, 0, 0, 0, _lib.stream_ptr())
        o = out8.cpu().numpy()
        assert abs(o[0] - float(fl)) / float(fl) < 1e-4, (o, float(fl))
        assert abs(o[1] - float(sl)) / float(sl) < 1e-4
        assert o[2] == (state == 1).sum()
        g = dcls.cpu().numpy()
        # clip boundary points, and p == 0.5 exactly where autograd's sub-gradient of max/abs at
        # z = 0 differs from the analytic derivative
        skip = np.zeros_like(g, bool); skip[0, :5, 0] = True
        assert rel_err(g[~skip], want_dlogit[~skip]) < 1e-4
        assert rel_err(dreg.cpu().numpy(), rt.grad.numpy()) < 1e-5


def test_losses_level_outputs_without_fp32_copies():
    """effdet_detection_losses with per-level bf16 gradient buffers: dcls_logits / dreg may be NULL (the
    tensor-core backward reads only the level buffers); losses and level buffers are bit-identical to the call
    that also writes the fp32 copies, and the level buffers hold bf16(fp32 gradient) at (image, cell, anchor*per+k)."""
    from efficientdet_b200 import _lib
    import ctypes
    lib = _lib.load()
    B, C, cells = 2, 8, [16, 4]
    N = 9 * sum(cells)
    cpad_cls, cpad_reg = 9 * C, 40
    rng = np.random.default_rng(3)
    p = rng.uniform(0.001, 0.999, (B, N, C)).astype(np.float32)
    reg = rng.normal(0, 1.2, (B, N, 4)).astype(np.float32)
    state = rng.choice([-1, 0, 1], (B, N), p=[0.1, 0.7, 0.2]).astype(np.float32)
    clsid = rng.integers(0, C, (B, N))
    reg_t = np.concatenate([rng.normal(0, 1, (B, N, 4)), state[..., None]], -1).astype(np.float32)
    d = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)
    pd, rd, rtd = d(p), d(reg), d(reg_t)
    st, cl = d(state, torch.int8), d(np.where(state == 1, clsid, -1), torch.int32)
    wsb = lib.effdet_detection_losses_workspace_size()
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    res = []
    for with_fp32 in (True, False):
        dcls = torch.zeros((B, N, C), device="cuda"); dreg = torch.zeros((B, N, 4), device="cuda")
        out8 = torch.zeros(8, device="cuda")
        lc = [torch.zeros((B, c, cpad_cls), device="cuda", dtype=torch.bfloat16) for c in cells]
        lr = [torch.zeros((B, c, cpad_reg), device="cuda", dtype=torch.bfloat16) for c in cells]
        pc = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in lc]); pr = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in lr])
        cc = (ctypes.c_int * 5)(*cells)
        _lib.call("effdet_detection_losses", pd.data_ptr(), rd.data_ptr(), rtd.data_ptr(), None, st.data_ptr(),
                  cl.data_ptr(), B, N, C, 0.25, 1.5, 1.0, 1.0, dcls.data_ptr() if with_fp32 else None,
                  dreg.data_ptr() if with_fp32 else None, out8.data_ptr(), ws.data_ptr(), wsb, pc, pr, cc,
                  len(cells), cpad_cls, cpad_reg, _lib.stream_ptr())
        torch.cuda.synchronize()
        res.append((out8.cpu(), [t.cpu() for t in lc], [t.cpu() for t in lr], dcls.cpu(), dreg.cpu()))
    a, b = res
    assert torch.equal(a[0], b[0])
    for x, y in zip(a[1] + a[2], b[1] + b[2]):
        assert torch.equal(x.view(torch.int16), y.view(torch.int16))
    assert float(b[3].abs().max()) == 0.0 and float(b[4].abs().max()) == 0.0      # untouched
    off = 0
    for l, c in enumerate(cells):
        want = a[3][:, off:off + 9 * c].reshape(B, c, 9 * C).to(torch.bfloat16)
        assert torch.equal(a[1][l][:, :, :9 * C].view(torch.int16), want.view(torch.int16))
        wantr = a[4][:, off:off + 9 * c].reshape(B, c, 36).to(torch.bfloat16)
        assert torch.equal(a[2][l][:, :, :36].view(torch.int16), wantr.view(torch.int16))
        off += 9 * c


def _d(a, dt=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)


def _nchw(a):
    return torch.from_numpy(np.ascontiguousarray(a)).permute(0, 3, 1, 2).contiguous().double()


@pytest.mark.parametrize("dtype,B,H,C,R", [("f32", 2, 12, 48, 4), ("bf16", 3, 9, 144, 6), ("bf16", 2, 8, 2688, 112),
                                            ("f32", 2, 6, 1152, 48)])
def test_fused_se_bn_backward_matches_separate_passes(dtype, B, H, C, R):
    """effdet_se_bn_backward (one reduction + one apply pass over dyg / z) against effdet_se_backward followed by
    effdet_bn_act_backward (swish), the pair the whole-step oracle tests validate: same dz, gamma / beta gradients
    and SE weight gradients (fp32 to 2e-5; bf16 to the rounding of the dy tensor the separate form stores)."""
    from efficientdet_b200 import _lib
    lib = _lib.load()
    dt = _lib.F32 if dtype == "f32" else _lib.BF16
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    HW = H * H
    g = torch.Generator(device="cuda").manual_seed(C + H)
    rnd = lambda *shape, s=1.0: torch.randn(*shape, device="cuda", generator=g) * s
    z = (rnd(B, HW, C) * 1.5 + rnd(C) * 0.5).to(tdt)
    dyg = (rnd(B, HW, C) * 0.1).to(tdt)
    gamma, beta = rnd(C) * 0.3 + 1.0, rnd(C) * 0.2
    zf = z.float().reshape(-1, C)
    mean, var = zf.mean(0), zf.var(0, unbiased=False)
    invstd = 1.0 / torch.sqrt(var + 1e-3)
    ua = (gamma * invstd).contiguous(); ub = (beta - mean * ua).contiguous()
    u = z.float() * ua + ub
    y = (u * torch.sigmoid(u)).to(tdt)
    w1, b1, w2, b2 = rnd(C, R, s=0.2), rnd(R, s=0.1), rnd(R, C, s=0.2), rnd(C, s=0.1)
    sblk = lib.effdet_se_backward_blocks(HW, C, dt)
    se_sum = torch.empty((B, sblk, C), device="cuda")
    st = _lib.stream_ptr()
    _lib.call("effdet_spatial_sum", y.data_ptr(), se_sum.data_ptr(), sblk, B, HW, C, dt, st)
    gate = torch.empty((B, C), device="cuda")
    _lib.call("effdet_se_gate", se_sum.data_ptr(), sblk, 1.0 / HW, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
              b2.data_ptr(), gate.data_ptr(), B, C, R, st)
    fcs_n = B * (2 * C * R + R + C)

    def grads():
        return [torch.zeros_like(t) for t in (w1, b1, w2, b2, gamma, beta)]
    # ---- separate passes
    dw1a, db1a, dw2a, db2a, dga, dba = grads()
    dy = torch.empty_like(z); dza = torch.empty_like(z)
    dgb = lib.effdet_se_backward_blocks(HW, C, dt)
    dgp = torch.empty(B * dgb * C, device="cuda"); fcs = torch.empty(fcs_n, device="cuda")
    dmean = torch.empty(B * C, device="cuda")
    _lib.call("effdet_se_backward", dyg.data_ptr(), y.data_ptr(), gate.data_ptr(), se_sum.data_ptr(), sblk,
              w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), dy.data_ptr(), dw1a.data_ptr(),
              db1a.data_ptr(), dw2a.data_ptr(), db2a.data_ptr(), dgp.data_ptr(), dgb, fcs.data_ptr(),
              dmean.data_ptr(), B, HW, C, R, dt, st)
    rows = B * HW
    nblk = lib.effdet_colreduce_blocks(rows, C, dt)
    part = torch.empty(2 * C * nblk, device="cuda"); k123 = torch.empty(3 * C, device="cuda")
    _lib.call("effdet_bn_act_backward", dy.data_ptr(), z.data_ptr(), rows, C, gamma.data_ptr(), mean.data_ptr(),
              invstd.data_ptr(), ua.data_ptr(), ub.data_ptr(), 0, _lib.ACT_SWISH, dga.data_ptr(), dba.data_ptr(),
              dza.data_ptr(), k123.data_ptr(), part.data_ptr(), nblk, dt, st)
    # ---- fused
    dw1b, db1b, dw2b, db2b, dgbb, dbb = grads()
    dzb = torch.empty_like(z)
    nb2 = lib.effdet_se_bn_backward_blocks(B, HW, C, dt)
    dgp2 = torch.empty(B * (nb2 + 1) * C, device="cuda"); bnp = torch.empty(B * (nb2 + 1) * 4 * C, device="cuda")
    bnr = torch.empty(B * 2 * C, device="cuda"); k2 = torch.empty(3 * C, device="cuda")
    fcs2 = torch.empty(fcs_n, device="cuda"); dmean2 = torch.empty(B * C, device="cuda")
    _lib.call("effdet_se_bn_backward", dyg.data_ptr(), z.data_ptr(), gate.data_ptr(), se_sum.data_ptr(), sblk,
              w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), dw1b.data_ptr(), db1b.data_ptr(),
              dw2b.data_ptr(), db2b.data_ptr(), gamma.data_ptr(), mean.data_ptr(), invstd.data_ptr(), ua.data_ptr(),
              ub.data_ptr(), dgbb.data_ptr(), dbb.data_ptr(), dzb.data_ptr(), k2.data_ptr(), dgp2.data_ptr(),
              bnp.data_ptr(), bnr.data_ptr(), nb2, fcs2.data_ptr(), dmean2.data_ptr(), B, HW, C, R, dt, st)
    torch.cuda.synchronize()
    tol = 2e-5 if dtype == "f32" else 1.5e-2
    pairs = [("dz", dza, dzb), ("dgamma", dga, dgbb), ("dbeta", dba, dbb), ("dw1", dw1a, dw1b), ("db1", db1a, db1b),
             ("dw2", dw2a, dw2b), ("db2", db2a, db2b), ("dmean", dmean, dmean2)]
    for name, a, b in pairs:
        a, b = a.float(), b.float()
        err = float((a - b).norm() / a.norm().clamp_min(1e-20))
        assert err < tol, (name, err)


# ------------------------------------------------------------------ backward kernels, one by one
def test_conv_wgrad_grouped_and_strided_dz():
    import ctypes
    from efficientdet_b200 import _lib
    from oracle import graph
    rng = np.random.default_rng(0)
    lib = _lib.load()
    for (k, stride, cin, cout, Hs, B) in [(3, 1, 64, 36, [8, 4, 2], 3), (1, 1, 40, 64, [6], 2),
                                           (3, 2, 64, 64, [8], 2), (3, 1, 24, 810, [4, 2], 2)]:
        w = torch.zeros((k, k, cin, cout), dtype=torch.float64, requires_grad=True)
        xs, dzs, total = [], [], 0
        for H in Hs:
            Ho = (H + stride - 1) // stride
            xs.append(rng.standard_normal((B, H, H, cin)).astype(np.float32))
            dzs.append(rng.standard_normal((B, Ho, Ho, cout)).astype(np.float32))
            y = graph.conv2d(_nchw(xs[-1]), w, stride)
            total = total + (y * _nchw(dzs[-1])).sum()
        total.backward()
        # dz packed like the concatenated head outputs: (B, sum HoWo, cout)
        cat = np.concatenate([z.reshape(B, -1, cout) for z in dzs], 1)
        catd = _d(cat)
        d = _lib.WgradDesc()
        d.n_groups = len(Hs)
        keep, off = [], 0
        for i, H in enumerate(Hs):
            xd = _d(xs[i]); keep.append(xd)
            Ho = (H + stride - 1) // stride
            d.x[i] = xd.data_ptr(); d.dz[i] = catd.data_ptr() + off * cout * 4
            d.H[i] = d.W[i] = H
            d.dz_ld[i] = cout; d.dz_batch_stride[i] = cat.shape[1] * cout
            off += Ho * Ho
        d.B, d.Cin, d.Cout, d.kh, d.kw, d.stride = B, cin, cout, k, k, stride
        ns = lib.effdet_conv_wgrad_splits(ctypes.byref(d))
        assert ns > 0
        part = torch.empty(ns * k * k * cin * cout, device="cuda")
        out = torch.full((k, k, cin, cout), float("nan"), device="cuda")
        d.dweight, d.partial, d.n_splits, d.accumulate = out.data_ptr(), part.data_ptr(), ns, 0
        d.x_dtype = d.dz_dtype = _lib.F32
        _lib.call("effdet_conv_wgrad", ctypes.byref(d), _lib.stream_ptr())
        assert rel_err(out.cpu().numpy(), w.grad.numpy()) < 1e-5, (k, stride, cin, cout)


def test_conv_dgrad_via_transposed_weights_and_strided():
    import ctypes
    from efficientdet_b200 import _lib
    from oracle import graph
    rng = np.random.default_rng(1)
    B = 3
    for (k, stride, cin, cout, H) in [(3, 1, 64, 64, 8), (1, 1, 64, 88, 5), (3, 2, 64, 64, 8),
                                      (3, 2, 320, 64, 4)]:
        w = (rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
        Ho = (H + stride - 1) // stride
        dz = rng.standard_normal((B, Ho, Ho, cout)).astype(np.float32)
        x = torch.tensor(rng.standard_normal((B, cin, H, H)), dtype=torch.float64, requires_grad=True)
        graph.conv2d(x, w.astype(np.float64), stride).backward(_nchw(dz))
        want = x.grad.permute(0, 2, 3, 1).numpy()
        wd, dzd = _d(w), _d(dz)
        prev = rng.standard_normal((B, H, H, cin)).astype(np.float32)
        mask = rng.standard_normal((B, H, H, cin)).astype(np.float32)
        if stride == 1:
            wt = torch.empty(k * k * cin * cout, device="cuda")
            _lib.call("effdet_conv_weight_transpose", wd.data_ptr(), wt.data_ptr(), k * k, cin, cout,
                      _lib.stream_ptr())
            out, md = _d(prev), _d(mask)
            d = _lib.ConvDesc()
            d.n_groups = 1
            d.x[0], d.y[0], d.residual[0], d.relu_mask[0] = dzd.data_ptr(), out.data_ptr(), out.data_ptr(), md.data_ptr()
            d.H[0] = d.W[0] = H
            d.B, d.Cin, d.Cout, d.kh, d.kw, d.stride = B, cout, cin, k, k, 1
            d.weight = wt.data_ptr()
            d.in_dtype = d.out_dtype = _lib.F32
            _lib.call("effdet_conv2d", ctypes.byref(d), _lib.stream_ptr())
            assert rel_err(out.cpu().numpy(), want * (mask > 0) + prev) < 1e-5
        else:
            out = _d(prev)
            _lib.call("effdet_conv_dgrad_strided", dzd.data_ptr(), wd.data_ptr(), out.data_ptr(), 1, B, H, H,
                      cin, cout, k, stride, _lib.F32, _lib.stream_ptr())
            assert rel_err(out.cpu().numpy(), want + prev) < 1e-5


@pytest.mark.parametrize("rows,C", [(4 * 12 * 12, 88), (140003, 64), (8 * 16 * 16, 224)])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_bn_train_forward_backward(dtype, rows, C):
    """(576, 88) and, in bf16, (2048, 224): tensors <= 512 KB take the one-launch thread-block-cluster backward;
    (140003, 64): the reduce / finalize / apply chain."""
    from efficientdet_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(2)
    tdt = torch.float32 if dtype == "fp32" else torch.bfloat16
    dt = _lib.F32 if dtype == "fp32" else _lib.BF16
    zd = _d(rng.standard_normal((rows, C)).astype(np.float32) * 1.5 + 0.3, tdt)
    dyd = _d(rng.standard_normal((rows, C)).astype(np.float32), tdt)
    gamma = rng.uniform(0.5, 1.5, C).astype(np.float32); beta = rng.normal(0, 0.2, C).astype(np.float32)
    mm0 = rng.normal(0, 0.1, C).astype(np.float32); mv0 = rng.uniform(0.5, 1.5, C).astype(np.float32)
    g_, b_, mm, mv = _d(gamma), _d(beta), _d(mm0), _d(mv0)
    sc, sh, mu, iv = (torch.empty(C, device="cuda") for _ in range(4))
    nblk = lib.effdet_colreduce_blocks(rows, C, dt)
    part = torch.empty(2 * C * nblk, device="cuda")
    _lib.call("effdet_bn_train_stats", zd.data_ptr(), rows, C, g_.data_ptr(), b_.data_ptr(), 1e-4, 0.997,
              mm.data_ptr(), mv.data_ptr(), sc.data_ptr(), sh.data_ptr(), mu.data_ptr(), iv.data_ptr(),
              part.data_ptr(), nblk, dt, _lib.stream_ptr())
    yd = torch.empty_like(zd)
    _lib.call("effdet_scale_shift_act", zd.data_ptr(), sc.data_ptr(), sh.data_ptr(), yd.data_ptr(), rows, C,
              _lib.ACT_RELU, dt, _lib.stream_ptr())
    z = zd.double().cpu().requires_grad_(True)
    gt, bt = torch.tensor(gamma, dtype=torch.float64, requires_grad=True), torch.tensor(beta, dtype=torch.float64, requires_grad=True)
    m, v = z.mean(0), z.var(0, unbiased=False)
    y = torch.relu((z - m) / torch.sqrt(v + 1e-4) * gt + bt)
    tol = 1e-5 if dtype == "fp32" else 1e-2
    assert rel_err(yd.double().cpu().numpy(), y.detach().numpy()) < tol
    assert rel_err(mm.cpu().numpy(), mm0 * 0.997 + m.detach().numpy() * 0.003) < 1e-5
    assert rel_err(mv.cpu().numpy(), mv0 * 0.997 + z.var(0, unbiased=True).detach().numpy() * 0.003) < 1e-5
    # backward uses the GPU's own y as the ReLU mask (no mask flips between implementations)
    ymask = (yd.double().cpu() > 0).double()
    ((z - m) / torch.sqrt(v + 1e-4) * gt + bt).backward(dyd.double().cpu() * ymask)
    dz = torch.empty_like(zd); k123 = torch.empty(3 * C, device="cuda")
    dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    _lib.call("effdet_bn_relu_backward", dyd.data_ptr(), yd.data_ptr(), zd.data_ptr(), rows, C, g_.data_ptr(),
              mu.data_ptr(), iv.data_ptr(), None, dg.data_ptr(), db.data_ptr(), dz.data_ptr(), k123.data_ptr(),
              part.data_ptr(), nblk, dt, _lib.stream_ptr())
    assert rel_err(dz.double().cpu().numpy(), z.grad.numpy()) < (1e-4 if dtype == "fp32" else 2e-2)
    assert rel_err(dg.cpu().numpy(), gt.grad.numpy()) < (1e-4 if dtype == "fp32" else 2e-2)
    assert rel_err(db.cpu().numpy(), bt.grad.numpy()) < (1e-4 if dtype == "fp32" else 2e-2)
    # frozen (inference-mode) BN: dz = scale * dy * [y > 0]
    _lib.call("effdet_bn_relu_backward", dyd.data_ptr(), yd.data_ptr(), zd.data_ptr(), rows, C, g_.data_ptr(),
              None, None, sc.data_ptr(), None, None, dz.data_ptr(), k123.data_ptr(), part.data_ptr(), 1, dt,
              _lib.stream_ptr())
    want = dyd.double().cpu() * ymask * sc.double().cpu()
    assert rel_err(dz.double().cpu().numpy(), want.numpy()) < (1e-6 if dtype == "fp32" else 1e-2)


@pytest.mark.parametrize("mode,three,weighted", [(1, False, True), (2, True, True), (2, False, False),
                                                 (1, False, False), (2, True, False)])
def test_fusion_and_depthwise_backward(mode, three, weighted):
    from efficientdet_b200 import _lib
    from oracle import graph
    lib = _lib.load()
    rng = np.random.default_rng(3 + mode)
    B, H, C = 2, 8, 88
    H0 = {1: H // 2, 2: H * 2}[mode]
    a = torch.tensor(rng.standard_normal((B, C, H0, H0)), dtype=torch.float64, requires_grad=True)
    b = torch.tensor(rng.standard_normal((B, C, H, H)), dtype=torch.float64, requires_grad=True)
    c = torch.tensor(rng.standard_normal((B, C, H, H)), dtype=torch.float64, requires_grad=True) if three else None
    fw = torch.tensor([0.7, -0.1, 0.4][:3 if three else 2], dtype=torch.float64, requires_grad=True)
    dw = torch.tensor(rng.standard_normal((3, 3, C, 1)) / 3, dtype=torch.float64, requires_grad=True)
    ra = graph.upsample2(a) if mode == 1 else graph.maxpool2(a)
    f = graph.fuse([ra, b] + ([c] if three else []), {"f/f": fw}, weighted, "f")
    f.retain_grad()
    z = graph.dwconv2d(f, dw, 1)
    dzn = rng.standard_normal((B, H, H, C)).astype(np.float32)
    z.backward(_nchw(dzn))
    nhwc = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().numpy().astype(np.float32)
    ad, bd, cd = _d(nhwc(a)), _d(nhwc(b)), (_d(nhwc(c)) if three else None)
    fwd_ = _d(fw.detach().numpy().astype(np.float32))
    fwp = fwd_.data_ptr() if weighted else None
    fd = torch.empty((B, H, H, C), device="cuda")
    _lib.call("effdet_resample_fuse", ad.data_ptr(), mode, bd.data_ptr(), cd.data_ptr() if three else None, fwp,
              1e-4, fd.data_ptr(), B, H, H, C, _lib.F32, _lib.stream_ptr())
    assert rel_err(fd.cpu().numpy(), nhwc(f)) < 1e-6
    dzd = _d(dzn)
    # depthwise weight gradient
    nblk = lib.effdet_dw_wgrad_blocks(B, H, H, C, _lib.F32)
    part = torch.empty(9 * C * nblk, device="cuda"); dk = torch.empty((3, 3, C), device="cuda")
    _lib.call("effdet_dw_wgrad", fd.data_ptr(), dzd.data_ptr(), B, H, H, C, dk.data_ptr(), part.data_ptr(), nblk,
              _lib.F32, _lib.stream_ptr())
    assert rel_err(dk.cpu().numpy(), dw.grad.numpy()[..., 0]) < 1e-5
    # depthwise data gradient = depthwise conv with the flipped kernel
    dwd = _d(dw.detach().numpy().astype(np.float32)); flip = torch.empty(9 * C, device="cuda")
    _lib.call("effdet_flip_taps", dwd.data_ptr(), flip.data_ptr(), 9, C, _lib.stream_ptr())
    ones, zeros = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    dfd = torch.empty((B, H, H, C), device="cuda")
    _lib.call("effdet_dwconv", dzd.data_ptr(), flip.data_ptr(), ones.data_ptr(), zeros.data_ptr(), dfd.data_ptr(),
              None, 0, B, H, H, C, 3, 1, _lib.ACT_NONE, _lib.F32, _lib.stream_ptr())
    assert rel_err(dfd.cpu().numpy(), nhwc(f.grad)) < 1e-5
    # fusion weights
    if weighted:
        fpart = torch.empty(4 * 148 * 8, device="cuda"); dfw = torch.empty(3, device="cuda")
        _lib.call("effdet_fuse_backward_weights", dfd.data_ptr(), fd.data_ptr(), ad.data_ptr(), mode, bd.data_ptr(),
                  cd.data_ptr() if three else None, fwp, 1e-4, dfw.data_ptr(), fpart.data_ptr(), B, H, H, C,
                  _lib.F32, _lib.stream_ptr())
        n_in = 3 if three else 2
        # with one active weight in_0 - f is a difference of nearly equal numbers: fp32 leaves
        # ~1e-3 relative accuracy in that (tiny) gradient
        assert rel_err(dfw.cpu().numpy()[:n_in], fw.grad.numpy()) < 2e-3
    # inputs (overwrite, then accumulate on top of existing content)
    n_in = 3 if three else 2
    for which, src in enumerate([a, b, c][:n_in]):
        want = nhwc(src.grad)
        dst = torch.full(want.shape, float("nan"), device="cuda")
        args = (dfd.data_ptr(), which, mode, ad.data_ptr(), fwp, n_in, 1e-4, dst.data_ptr())
        _lib.call("effdet_fuse_backward_input", *args, 0, B, H, H, C, _lib.F32, _lib.stream_ptr())
        assert rel_err(dst.cpu().numpy(), want) < 1e-5, which
        _lib.call("effdet_fuse_backward_input", *args, 1, B, H, H, C, _lib.F32, _lib.stream_ptr())
        assert rel_err(dst.cpu().numpy(), 2 * want) < 1e-5, which


def test_colsum_fold_and_sgd():
    from efficientdet_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(4)
    x = rng.standard_normal((600, 810)).astype(np.float32)          # 810 % 4 != 0 -> fold 2
    xd = _d(x)
    nblk = lib.effdet_colreduce_blocks(300, 1620, _lib.F32)
    part = torch.empty(2 * 1620 * nblk, device="cuda"); out = torch.ones(810, device="cuda")
    _lib.call("effdet_colsum", xd.data_ptr(), 300, 1620, 2, out.data_ptr(), 1, part.data_ptr(), nblk, _lib.F32,
              _lib.stream_ptr())
    assert rel_err(out.cpu().numpy(), 1 + x.astype(np.float64).sum(0)) < 1e-5
    n = 100003
    w, g, v = (rng.standard_normal(n).astype(np.float32) for _ in range(3))
    wd, gd, vd = _d(w), _d(g), _d(v)
    _lib.call("effdet_sgd_momentum_step", wd.data_ptr(), gd.data_ptr(), vd.data_ptr(), n, 0.01 / (1 + 4e-5 * 7), 0.9,
              0.5, _lib.stream_ptr())
    from oracle import losses
    wt, vt = torch.tensor(w, dtype=torch.float64), torch.tensor(v, dtype=torch.float64)
    losses.sgd_momentum_step(wt, torch.tensor(g, dtype=torch.float64) * 0.5, vt, 0.01, 4e-5, 0.9, 7)
    assert rel_err(wd.cpu().numpy(), wt.numpy()) < 1e-6 and rel_err(vd.cpu().numpy(), vt.numpy()) < 1e-6


# ------------------------------------------------------------------ whole step
@pytest.mark.parametrize("weighted,freeze_bn", [(False, False), (True, False), (True, True)])
def test_training_step_gradients(weighted, freeze_bn):
    """End to end against the fp64 autograd oracle.  Per-tensor relative L2 <= 5e-2: the oracle
    itself moves by up to ~1e-2 (L2) between fp32 and fp64 on this network because a handful of the
    ~10^7 ReLU units sit within rounding distance of zero and the detection gradients are carried
    by few positive anchors (measured with the oracle alone, see tests/test_gpu_whole_model.py); the kernels themselves are pinned
    much tighter by the single-kernel tests above, which share the ReLU masks."""
    from efficientdet_b200.model import efficientdet, EFFICIENTNET_DEPTHS
    from efficientdet_b200.optimizers import SGD
    from efficientdet_b200.utils.tpu import tpu_focal, tpu_smooth_l1
    from oracle import train as otrain
    from util_model import rel_l2
    size, C, B, phi = 256, 5, 8, 0
    model = efficientdet(phi, num_classes=C, weighted_bifpn=weighted, freeze_bn=freeze_bn,
                         image_size=size, dtype="fp32", drop_connect_rate=0, just_training_model=True)
    W0 = perturb_weights(model)
    assert EFFICIENTNET_DEPTHS[phi] == 227
    for i in range(1, model.backbone_depth):     # == EFFICIENTNET_DEPTHS minus the *_drop layers
        model.layers[i].trainable = False
    model.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9),
                  loss={"regression": tpu_smooth_l1(), "classification": tpu_focal(alpha=0.25, gamma=1.5)})
    anchors, ann, reg_t, lab_t = _targets(size, B, C)
    rng = np.random.default_rng(5)
    img = rng.standard_normal((B, size, size, 3)).astype(np.float32)
    total, l_reg, l_cls = model.train_on_batch(img, [reg_t, lab_t])
    fl, sl, grads, stats = otrain.loss_and_grads(W0, img, reg_t, lab_t, phi, C, weighted, freeze_bn)
    assert abs(l_cls - fl) / fl < 1e-4, (l_cls, fl)
    assert abs(l_reg - sl) / max(sl, 1e-9) < 1e-4, (l_reg, sl)
    net = model.net
    bad = {}
    for k, g in grads.items():
        got = net.grads[k].cpu().numpy()
        if np.abs(g).max() < 1e-12:
            continue
        e = rel_l2(got, g)
        # fusion-weight gradients are global sums with heavy cancellation: looser bound
        if not e < (0.2 if k.startswith("w_bi_fpn_add") else 5e-2):
            bad[k] = float(e)
    assert not bad, bad
    # SGD: first step from zero velocity: w1 = w0 - lr * g  (checked with the GPU's own gradients)
    W1 = model.get_weights_dict()
    for k in list(grads)[:60]:
        want = W0[k].astype(np.float64) - 0.01 * net.grads[k].cpu().numpy()
        assert rel_err(W1[k], want) < 1e-5, k
    # BN moving averages follow momentum .997 with the batch statistics (training-mode BN only)
    if not freeze_bn:
        for name, (m, v) in list(stats.items())[:12]:
            want_m = W0[name + "/moving_mean"] * 0.997 + m * 0.003
            want_v = W0[name + "/moving_variance"] * 0.997 + v * 0.003
            assert rel_err(W1[name + "/moving_mean"], want_m) < 1e-4, name
            assert rel_err(W1[name + "/moving_variance"], want_v) < 1e-4, name
    else:
        assert np.array_equal(W1["BiFPN_0_P3_bn/gamma"], W0["BiFPN_0_P3_bn/gamma"])
    # backbone untouched
    assert np.array_equal(W1["stem_conv/kernel"], W0["stem_conv/kernel"])
    # compact device targets give the same step (bit-identical: the step is deterministic)
    model2 = efficientdet(phi, num_classes=C, weighted_bifpn=weighted, freeze_bn=freeze_bn,
                          image_size=size, dtype="fp32", drop_connect_rate=0, just_training_model=True)
    model2.set_weights_dict(W0)
    model2.freeze_backbone()
    model2.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9))
    from efficientdet_b200.utils.anchors import anchor_targets_device
    r, _, st, cl = anchor_targets_device(anchors, [(size, size, 3)] * B, ann, C, dense_labels=False,
                                         compact=True)
    t2 = model2.train_on_batch(img, (r, st, cl))
    assert abs(t2[0] - total) / total < 1e-6
    k = "class_head/pyramid_classification/kernel"
    assert np.array_equal(model2.net.grads[k].cpu().numpy(), net.grads[k].cpu().numpy())


def test_training_step_bf16_tensor_cores_matches_simt():
    """bf16 speed mode: the tcgen05 convolution path and the SIMT path agree on losses and
    gradients up to the noise floor of bf16 training on this problem.  The floor is measured, not
    assumed: a third (SIMT) run on inputs perturbed by 1e-3 relative noise moves the gradients by
    20-60 % (relative L2) in the deep BiFPN layers, because activations and activation gradients
    are stored in bf16; the tensor-core run must not move them more than 2x that (+0.1)."""
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.optimizers import SGD
    from util_model import rel_l2
    size, C, B, phi = 256, 5, 4, 0
    anchors, ann, reg_t, lab_t = _targets(size, B, C)
    img = np.random.default_rng(5).standard_normal((B, size, size, 3)).astype(np.float32)
    img_noisy = img * (1 + 1e-3 * np.random.default_rng(1).standard_normal(img.shape).astype(np.float32))
    res = {}
    for tag, tc, x in (("simt", False, img), ("tc", True, img), ("noise", False, img_noisy)):
        model = efficientdet(phi, num_classes=C, weighted_bifpn=True, image_size=size, dtype="bf16",
                             drop_connect_rate=0, just_training_model=True, tensor_cores=tc)
        perturb_weights(model)
        model.freeze_backbone()
        model.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9))
        loss = model.train_on_batch(x, [reg_t, lab_t])
        kinds = {op.kind for op in list(model._trainer.plans.values())[0].ops}
        assert any(k.endswith("_tc") for k in kinds) == tc
        res[tag] = (loss, {k: v.cpu().numpy().copy() for k, v in model.net.grads.items()
                           if k.startswith(("BiFPN_", "box_head", "class_head"))})
    assert abs(res["tc"][0][0] - res["simt"][0][0]) / res["simt"][0][0] < 1e-2
    bad = {}
    for k, g in res["simt"][1].items():
        if k.endswith(("moving_mean", "moving_variance")) or np.abs(g).max() < 1e-10:
            continue
        e, floor = rel_l2(res["tc"][1][k], g), rel_l2(res["noise"][1][k], g)
        if not e < 2.0 * floor + 0.1:
            bad[k] = (float(e), float(floor))
    assert not bad, bad


@pytest.mark.parametrize("weighted,freeze_bn", [(True, False), (False, True)])
def test_full_training_step_with_backbone(weighted, freeze_bn):
    """Nothing frozen (BASELINE config 4 shape of step): gradients of EVERY weight, backbone included
    (MBConv expand / depthwise / squeeze-excite / project, stem), against the fp64 autograd oracle."""
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.optimizers import SGD
    from oracle import train as otrain
    from util_model import rel_l2
    size, C, B, phi = 256, 5, 8, 0
    model = efficientdet(phi, num_classes=C, weighted_bifpn=weighted, freeze_bn=freeze_bn, image_size=size,
                         dtype="fp32", drop_connect_rate=0, just_training_model=True)
    W0 = perturb_weights(model)
    model.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9))
    anchors, ann, reg_t, lab_t = _targets(size, B, C)
    img = np.random.default_rng(5).standard_normal((B, size, size, 3)).astype(np.float32)
    total, l_reg, l_cls = model.train_on_batch(img, [reg_t, lab_t])
    fl, sl, grads, stats = otrain.loss_and_grads(W0, img, reg_t, lab_t, phi, C, weighted, freeze_bn,
                                                 freeze_backbone=False)
    assert abs(l_cls - fl) / fl < 2e-4, (l_cls, fl)
    assert abs(l_reg - sl) / max(sl, 1e-9) < 2e-4, (l_reg, sl)
    assert any(k.startswith("block") for k in grads) and "stem_conv/kernel" in grads
    net = model.net
    bad = {}
    for k, g in grads.items():
        if np.abs(g).max() < 1e-12:
            continue
        e = rel_l2(net.grads[k].cpu().numpy(), g)
        if not e < (0.25 if k.startswith("w_bi_fpn_add") else 8e-2):
            bad[k] = float(e)
    assert not bad, bad
    W1 = model.get_weights_dict()
    for k in ("stem_conv/kernel", "block3b_se_reduce/kernel", "block5a_dwconv/depthwise_kernel"):
        want = W0[k].astype(np.float64) - 0.01 * net.grads[k].cpu().numpy()
        assert rel_err(W1[k], want) < 1e-5, k
    if not freeze_bn:
        for name in ("stem_bn", "block2a_expand_bn", "block4b_bn", "block7a_project_bn"):
            m, v = stats[name]
            assert rel_err(W1[name + "/moving_mean"], W0[name + "/moving_mean"] * 0.99 + m * 0.01) < 1e-4, name
            assert rel_err(W1[name + "/moving_variance"], W0[name + "/moving_variance"] * 0.99 + v * 0.01) < 1e-4


def test_full_training_step_bf16_against_fp32():
    """bf16 speed mode with nothing frozen (BASELINE config 4's step): the TMA / tcgen05 kernels of the
    backbone backward (depthwise, SE, BN+swish, stem) against the fp32 accuracy mode of the same step.
    On this random-init problem the gradient DIRECTION of deep layers is chaotic under bf16 activation
    rounding (ReLU masks at rounding distance from zero, few positive anchors: relative L2 differences
    of 0.7-0.9 are measured even in the BiFPN, see the noise-floor test above), so only the loss and
    the gradient magnitudes are compared here; the kernels themselves are checked tightly against
    autograd in the kernel-level tests and, through the fp32 mode, in the test above."""
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.optimizers import SGD
    from util_model import rel_l2
    size, C, B, phi = 256, 5, 4, 0
    anchors, ann, reg_t, lab_t = _targets(size, B, C)
    img = np.random.default_rng(5).standard_normal((B, size, size, 3)).astype(np.float32)
    res = {}
    for dt in ("fp32", "bf16"):
        model = efficientdet(phi, num_classes=C, weighted_bifpn=False, image_size=size, dtype=dt,
                             drop_connect_rate=0, just_training_model=True)
        perturb_weights(model)
        model.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9))
        loss = model.train_on_batch(img, [reg_t, lab_t])
        res[dt] = (loss, {k: v.cpu().numpy().copy() for k, v in model.net.grads.items()})
    assert abs(res["bf16"][0][0] - res["fp32"][0][0]) / res["fp32"][0][0] < 3e-2
    keys = ["stem_conv/kernel", "block1a_dwconv/depthwise_kernel", "block2a_expand_conv/kernel",
            "block3a_dwconv/depthwise_kernel", "block5b_se_reduce/kernel", "block5b_se_expand/kernel",
            "block6a_dwconv/depthwise_kernel", "block7a_project_conv/kernel", "block4b_bn/gamma"]
    bad = {}
    for k in keys:
        a, b = np.linalg.norm(res["bf16"][1][k]), np.linalg.norm(res["fp32"][1][k])
        if not (np.isfinite(a) and 0.6 < a / b < 1.6):
            bad[k] = (float(a), float(b))
    assert not bad, bad


@pytest.mark.parametrize("k,stride,C,H,B", [(3, 1, 48, 20, 3), (5, 1, 144, 17, 2), (3, 2, 96, 16, 2), (5, 2, 240, 18, 2),
                                            (5, 1, 64, 8, 5), (3, 1, 32, 33, 1)])
def test_depthwise_backward_bf16_tma(k, stride, C, H, B):
    """bf16 depthwise backward (efficientnet.py:242-252): TMA-tiled weight gradient and the data gradient
    (forward kernel on dz with reversed taps for stride 1, gather form for stride 2) against autograd
    on the same bf16-rounded operands."""
    from efficientdet_b200 import _lib
    from oracle import graph
    lib = _lib.load()
    rng = np.random.default_rng(k * 100 + C)
    Ho = (H + stride - 1) // stride
    x = torch.from_numpy(rng.standard_normal((B, H, H, C)).astype(np.float32)).cuda().to(torch.bfloat16)
    dz = torch.from_numpy(rng.standard_normal((B, Ho, Ho, C)).astype(np.float32)).cuda().to(torch.bfloat16)
    w = torch.from_numpy((rng.standard_normal((k, k, C)) * 0.3).astype(np.float32)).cuda()
    dx = torch.full((B, H, H, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros((k, k, C), device="cuda")
    nb = lib.effdet_dw_backward_blocks(B, H, H, C, k, stride, _lib.BF16)
    part = torch.empty(k * k * C * nb, device="cuda")
    _lib.call("effdet_dw_backward", x.data_ptr(), dz.data_ptr(), w.data_ptr(), dx.data_ptr(), dw.data_ptr(),
              part.data_ptr(), nb, B, H, H, C, k, stride, _lib.BF16, _lib.stream_ptr())
    torch.cuda.synchronize()
    xr = x.float().cpu().double().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.cpu().double().requires_grad_(True)
    wt = wr.permute(2, 0, 1)[:, None]
    y = torch.nn.functional.conv2d(graph.same_pad(xr, k, stride), wt, None, stride=stride, groups=C)
    y.backward(dz.float().cpu().double().permute(0, 3, 1, 2))
    assert rel_err(dw.cpu().numpy(), wr.grad.numpy()) < 2e-5
    assert rel_err(dx.float().cpu().numpy(), xr.grad.permute(0, 2, 3, 1).numpy()) < 6e-3


@pytest.mark.parametrize("k,C,H,B", [(3, 96, 16, 2), (5, 240, 18, 2), (3, 48, 21, 1), (5, 64, 33, 2)])
def test_depthwise_stride2_dgrad_by_zero_insertion(k, C, H, B):
    """The training plan's stride-2 depthwise data gradient (TrainPlan._dw_dgrad_stride2): dz zero-inserted
    at offset (k-1)/2 - pad_t, then the stride-1 TMA depthwise kernel with flipped taps -- against autograd
    of the stride-2 SAME depthwise conv (efficientnet.py:242-252), even and odd input sizes."""
    from efficientdet_b200 import _lib
    from oracle import graph
    rng = np.random.default_rng(k * 10 + C + H)
    Ho = (H + 1) // 2
    dz = torch.from_numpy(rng.standard_normal((B, Ho, Ho, C)).astype(np.float32)).cuda().to(torch.bfloat16)
    w = torch.from_numpy((rng.standard_normal((k, k, C)) * 0.3).astype(np.float32)).cuda()
    pad_t = max((Ho - 1) * 2 + k - H, 0) // 2
    shift = (k - 1) // 2 - pad_t
    U = torch.full((B, H, H, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    wflip = torch.empty_like(w)
    dx = torch.full((B, H, H, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    ones, zeros = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    st = _lib.stream_ptr()
    _lib.call("effdet_zero_insert", dz.data_ptr(), U.data_ptr(), B, Ho, Ho, C, H, H, shift, shift, _lib.BF16, st)
    _lib.call("effdet_flip_taps", w.data_ptr(), wflip.data_ptr(), k * k, C, st)
    _lib.call("effdet_dwconv", U.data_ptr(), wflip.data_ptr(), ones.data_ptr(), zeros.data_ptr(), dx.data_ptr(),
              None, 0, B, H, H, C, k, 1, _lib.ACT_NONE, _lib.BF16, st)
    torch.cuda.synchronize()
    xr = torch.zeros((B, C, H, H), dtype=torch.float64, requires_grad=True)
    wt = w.cpu().double().permute(2, 0, 1)[:, None]
    y = torch.nn.functional.conv2d(graph.same_pad(xr, k, 2), wt, None, stride=2, groups=C)
    y.backward(dz.float().cpu().double().permute(0, 3, 1, 2))
    assert rel_err(dx.float().cpu().numpy(), xr.grad.permute(0, 2, 3, 1).numpy()) < 6e-3


@pytest.mark.parametrize("C0,H,B", [(32, 64, 2), (48, 70, 3), (40, 33, 1)])
def test_stem_weight_gradient_bf16(C0, H, B):
    """Stem conv (efficientnet.py:413-423: 3x3, stride 2, 3 -> C0) weight gradient from bf16 dz (tiled
    kernel) against autograd on the same operands."""
    from efficientdet_b200 import _lib
    from oracle import graph
    lib = _lib.load()
    rng = np.random.default_rng(C0 + H)
    Ho = (H + 1) // 2
    img = torch.from_numpy(rng.standard_normal((B, H, H, 3)).astype(np.float32)).cuda()
    dz = torch.from_numpy(rng.standard_normal((B, Ho, Ho, C0)).astype(np.float32)).cuda().to(torch.bfloat16)
    dw = torch.zeros((3, 3, 3, C0), device="cuda")
    nb = lib.effdet_stem_wgrad_blocks(B, H, H)
    part = torch.empty(27 * C0 * nb, device="cuda")
    _lib.call("effdet_stem_wgrad", img.data_ptr(), dz.data_ptr(), dw.data_ptr(), part.data_ptr(), nb, B, H, H, C0,
              _lib.BF16, _lib.stream_ptr())
    torch.cuda.synchronize()
    w = torch.zeros((C0, 3, 3, 3), dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(graph.same_pad(img.cpu().double().permute(0, 3, 1, 2), 3, 2), w, None, stride=2)
    y.backward(dz.float().cpu().double().permute(0, 3, 1, 2))
    assert rel_err(dw.cpu().numpy(), w.grad.permute(2, 3, 1, 0).numpy()) < 2e-5


def test_partial_backbone_freeze_is_rejected():
    from efficientdet_b200.model import efficientdet
    z = lambda *s: np.zeros(s, np.float32)
    model = efficientdet(0, num_classes=3, image_size=128, just_training_model=True, drop_connect_rate=0)
    model.layers[5].trainable = False
    model.compile()
    with pytest.raises(NotImplementedError):
        model.train_on_batch(z(1, 128, 128, 3), [z(1, 3069, 5), z(1, 3069, 4)])


def test_drop_connect_scales_distribution_and_fresh_draws():
    """effdet_drop_connect_scales: Keras dropout semantics (keep = u >= rate, kept branch * 1/(1-rate)),
    a new draw per launch (device-side step counter), reproducible from (seed, step)."""
    from efficientdet_b200 import _lib
    DEV = torch.device("cuda")
    lib = _lib.load()
    nb, B = 9, 4096
    rates = torch.tensor([0.0125 * (i + 1) for i in range(nb)], dtype=torch.float32, device=DEV)
    step = torch.zeros(1, dtype=torch.int64, device=DEV)
    a, b = torch.empty((nb, B), device=DEV), torch.empty((nb, B), device=DEV)
    _lib.check(lib.effdet_drop_connect_scales(rates.data_ptr(), nb, B, 77, step.data_ptr(), a.data_ptr(), None))
    _lib.check(lib.effdet_drop_connect_scales(rates.data_ptr(), nb, B, 77, step.data_ptr(), b.data_ptr(), None))
    assert int(step.item()) == 2
    a, b, r = a.cpu().numpy(), b.cpu().numpy(), rates.cpu().numpy()
    assert not np.array_equal(a, b)
    for i in range(nb):
        keep = np.float32(1.0) / (np.float32(1.0) - r[i])
        assert set(np.unique(a[i])) <= {np.float32(0.0), keep}
        frac = float((a[i] == 0).mean())
        assert abs(frac - r[i]) < 4 * np.sqrt(r[i] * (1 - r[i]) / B) + 1e-3, (i, frac, r[i])
    step.zero_()
    c = torch.empty((nb, B), device=DEV)
    _lib.check(lib.effdet_drop_connect_scales(rates.data_ptr(), nb, B, 77, step.data_ptr(), c.data_ptr(), None))
    assert np.array_equal(c.cpu().numpy(), a)


@pytest.mark.parametrize("freeze_backbone", [False, True])
def test_training_step_with_stochastic_depth(freeze_backbone):
    """drop_connect_rate = 0.2 (the reference default, efficientnet.py:318): the training step draws a
    per-(block, image) keep mask on the device; with the SAME mask the fp64 oracle gives the same losses
    and gradients (forward scale + residual, backward through the dropped branch), and a second step
    draws a different mask.  Frozen backbone: Keras dropout ignores `trainable`, so the mask still acts."""
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.optimizers import SGD
    from oracle import train as otrain
    from util_model import rel_l2
    size, C, B, phi = 256, 5, 8, 0
    model = efficientdet(phi, num_classes=C, weighted_bifpn=False, image_size=size, dtype="fp32",
                         drop_connect_rate=0.6, just_training_model=True)     # high rate: drops do happen at B=8
    W0 = perturb_weights(model)
    if freeze_backbone:
        model.freeze_backbone()
    model.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9))
    anchors, ann, reg_t, lab_t = _targets(size, B, C)
    img = np.random.default_rng(5).standard_normal((B, size, size, 3)).astype(np.float32)
    total, l_reg, l_cls = model.train_on_batch(img, [reg_t, lab_t])
    plan = model._trainer.plan(B, dense=True)
    assert len(plan.drop_blocks) == 9            # B0: 16 blocks, 9 of them skip blocks
    scales = plan.drop_scales.cpu().numpy().copy()
    drop = {b.prefix: scales[i] for i, b in enumerate(plan.drop_blocks)}
    for i, b in enumerate(plan.drop_blocks):
        keep = np.float32(1.0) / (np.float32(1.0) - np.float32(b.drop_rate))
        assert set(np.unique(scales[i])) <= {np.float32(0.0), keep}
    assert (scales == 0).any() and (scales != 0).any()
    fl, sl, grads, _ = otrain.loss_and_grads(W0, img, reg_t, lab_t, phi, C, False, False,
                                             freeze_backbone=freeze_backbone, drop_scale=drop)
    assert abs(l_cls - fl) / fl < 2e-4, (l_cls, fl)
    assert abs(l_reg - sl) / max(sl, 1e-9) < 2e-4, (l_reg, sl)
    bad = {}
    for k, g in grads.items():
        if np.abs(g).max() < 1e-12:
            continue
        e = rel_l2(model.net.grads[k].cpu().numpy(), g)
        if not e < 8e-2:
            bad[k] = float(e)
    assert not bad, bad
    model.train_on_batch(img, [reg_t, lab_t])
    assert not np.array_equal(plan.drop_scales.cpu().numpy(), scales)
    # inference is unaffected by the mask (FixedDropout only acts in the training phase)
    assert model.net.drop_scale == {}


@pytest.mark.parametrize("freeze_backbone,dtype", [(True, "bf16"), (False, "bf16"), (False, "fp32")])
def test_training_step_on_raw_uint8_images(freeze_backbone, dtype):
    """train_on_batch(uint8 letterboxed images) == train_on_batch(normalize_image(images)): the TFRecord path
    of the reference decodes uint8 PNGs and normalises them on the fly (train_tpu.py:170-183); here the
    normalisation runs inside the stem (frozen backbone) or once on the device (trainable stem, whose weight
    gradient needs the float image).  The step is deterministic, so losses and gradients are bit-identical."""
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.optimizers import SGD
    from efficientdet_b200.utils.preprocess import normalize_image
    size, C, B = 128, 4, 4
    anchors, ann, reg_t, lab_t = _targets(size, B, C)
    img8 = np.random.default_rng(11).integers(0, 256, (B, size, size, 3), dtype=np.uint8)
    res = []
    for images in (img8, normalize_image(img8)):
        model = efficientdet(0, num_classes=C, image_size=size, dtype=dtype, drop_connect_rate=0,
                             just_training_model=True, seed=5)
        if freeze_backbone:
            model.freeze_backbone()
        model.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9))
        out = model.train_on_batch(images, [reg_t, lab_t])
        g = model.net.grads
        res.append((out, g["class_head/pyramid_classification/kernel"].cpu().numpy().copy(),
                    g["stem_conv/kernel"].cpu().numpy().copy()))
    assert res[0][0] == res[1][0]
    assert np.array_equal(res[0][1], res[1][1])
    if not freeze_backbone:
        assert np.abs(res[0][2]).max() > 0 and np.array_equal(res[0][2], res[1][2])


def test_fit_prefetched_matches_stepwise_training():
    """Trainer.fit_prefetched (copy stream + device target assignment + captured graph) produces the same
    losses, step after step, as train_on_batch with host-computed dense targets (fp32 mode)."""
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.optimizers import SGD
    from efficientdet_b200.utils.anchors import _pack_annotations, anchors_for_shape
    from oracle import anchors as oa
    size, C, B, phi, steps = 128, 4, 2, 0, 3
    anchors = anchors_for_shape((size, size))
    rng = np.random.default_rng(11)
    data = []
    for s in range(steps):
        ann = []
        for _ in range(B):
            n = int(rng.integers(1, 4))
            wh = rng.uniform(16, 60, (n, 2)); xy = rng.uniform(0, size - 64, (n, 2))
            ann.append({"bboxes": np.concatenate([xy, xy + wh], 1).astype(np.float32),
                        "labels": rng.integers(0, C, n).astype(np.float32)})
        data.append((rng.standard_normal((B, size, size, 3)).astype(np.float32), ann))
    losses = {}
    for mode in ("stepwise", "prefetched"):
        model = efficientdet(phi, num_classes=C, image_size=size, dtype="fp32", drop_connect_rate=0,
                             just_training_model=True)
        perturb_weights(model)
        model.freeze_backbone()
        model.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9))
        if mode == "stepwise":
            out = []
            for img, ann in data:
                reg_t, lab_t = oa.anchor_targets_bbox(anchors, [(size, size, 3)] * B, ann, C)
                tot, l_reg, l_cls = model.train_on_batch(img, [reg_t, lab_t])
                out.append((l_cls, l_reg))
        else:
            tr = model._trainer
            plan = tr.plan(B, dense=False)
            anchors_d = torch.from_numpy(anchors).cuda()
            def gen():
                for img, ann in data:
                    gt, gl, cnt, kmax = _pack_annotations(ann)
                    hw = np.tile(np.array([[float(size), float(size)]]), (B, 1))
                    yield (torch.from_numpy(img).pin_memory(),
                           [torch.from_numpy(a).pin_memory() for a in (gt, gl, cnt, hw)], kmax)
            out = [(float(o[0]), float(o[1])) for o in tr.fit_prefetched(plan, anchors_d, gen())]
        losses[mode] = out
    for a, b in zip(losses["stepwise"], losses["prefetched"]):
        assert abs(a[0] - b[0]) / max(abs(a[0]), 1e-9) < 1e-4 and abs(a[1] - b[1]) / max(abs(a[1]), 1e-9) < 1e-4, losses


def test_checkpoint_saver_round_trip_and_lr_schedule_in_fit(tmp_path):
    """utils/train.py:66-92 CheckpointSaver + utils/lr_schedule.py callback driven by Model.fit: the
    checkpoint restores weights, SGD velocity and the iteration count exactly."""
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.optimizers import SGD
    from efficientdet_b200.utils.lr_schedule import get_cosine_decay_with_linear_warmup
    from efficientdet_b200.utils.train import CheckpointSaver, restore_checkpoint
    size, C, B = 128, 3, 2
    anchors, ann, reg_t, lab_t = _targets(size, B, C)
    img = np.random.default_rng(2).standard_normal((B, size, size, 3)).astype(np.float32)

    def build():
        m = efficientdet(0, num_classes=C, image_size=size, dtype="fp32", drop_connect_rate=0,
                         just_training_model=True)
        perturb_weights(m)
        m.freeze_backbone()
        m.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9))
        return m
    m1 = build()
    sched = get_cosine_decay_with_linear_warmup(total_epochs=4, learning_rate_max=0.02, warmup_percent=0.5)
    saver = CheckpointSaver(str(tmp_path / "ckpt_{epoch:02d}_{loss:.3f}"))
    hist = m1.fit([(img, [reg_t, lab_t])] * 2, epochs=2, callbacks=[sched, saver])
    assert len(hist) == 4 and m1.optimizer.lr == pytest.approx(0.02) and m1.optimizer.iterations == 4
    files = sorted(tmp_path.glob("ckpt_01_*.npz"))
    assert len(files) == 1
    m2 = build()
    restore_checkpoint(m2, str(tmp_path))                  # newest checkpoint of the directory
    assert m2.optimizer.iterations == 4 and m2.optimizer.lr == pytest.approx(0.02)
    w1, w2 = m1.get_weights_dict(), m2.get_weights_dict()
    assert all(np.array_equal(w1[k], w2[k]) for k in w1)
    assert torch.equal(m1.net.velocity, m2.net.velocity)
    a, b = m1.train_on_batch(img, [reg_t, lab_t]), m2.train_on_batch(img, [reg_t, lab_t])
    assert a == b                                          # the restored model continues identically
