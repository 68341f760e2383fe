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
                  _lib.stream_ptr())
        o = out8.cpu().numpy()
        assert abs(o[0] - float(fl)) / float(fl) < 1e-4, (o, float(fl))
        assert abs(o[1] - float(sl)) / float(sl) < 1e-4
        assert o[2] == (state == 1).sum()
        g = dcls.cpu().numpy()
        skip = np.zeros_like(g, bool); skip[0, :4, 0] = True      # clip boundary points
        assert rel_err(g[~skip], want_dlogit[~skip]) < 1e-4
        assert rel_err(dreg.cpu().numpy(), rt.grad.numpy()) < 1e-5


@pytest.mark.parametrize("weighted,freeze_bn", [(False, False), (True, False), (True, True)])
def test_training_step_gradients(weighted, freeze_bn):
    from efficientdet_b200.model import efficientdet, EFFICIENTNET_DEPTHS
    from efficientdet_b200.optimizers import SGD
    from efficientdet_b200.utils.tpu import tpu_focal, tpu_smooth_l1
    from oracle import train as otrain
    size, C, B, phi = 128, 5, 4, 0
    model = efficientdet(phi, num_classes=C, weighted_bifpn=weighted, freeze_bn=freeze_bn,
                         image_size=size, dtype="fp32", drop_connect_rate=0, just_training_model=True)
    W0 = perturb_weights(model)
    for i in range(1, EFFICIENTNET_DEPTHS[phi]):
        model.layers[i].trainable = False
    model.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9),
                  loss={"regression": tpu_smooth_l1(), "classification": tpu_focal(alpha=0.25, gamma=1.5)})
    anchors, ann, reg_t, lab_t = _targets(size, B, C)
    rng = np.random.default_rng(5)
    img = rng.standard_normal((B, size, size, 3)).astype(np.float32)
    total, l_reg, l_cls = model.train_on_batch(img, [reg_t, lab_t])
    fl, sl, grads, stats = otrain.loss_and_grads(W0, img, reg_t, lab_t, phi, C, weighted, freeze_bn)
    assert abs(l_cls - fl) / fl < 1e-3, (l_cls, fl)
    assert abs(l_reg - sl) / max(sl, 1e-9) < 1e-3, (l_reg, sl)
    net = model.net
    bad = {}
    for k, g in grads.items():
        got = net.grads[k].cpu().numpy()
        scale = max(np.abs(g).max(), 1e-12)
        e = np.abs(got - g).max() / scale
        if not e < 2e-3:
            bad[k] = (float(e), float(scale))
    assert not bad, bad
    # SGD: w1 = w0 + v, v = -lr*g (first step, zero velocity)
    W1 = model.get_weights_dict()
    for k in list(grads)[:50]:
        want = W0[k] - 0.01 * grads[k]
        assert rel_err(W1[k], want) < 1e-4, k
    # BN moving averages follow momentum .997 with the batch statistics (training-mode BN only)
    if not freeze_bn:
        for name, (m, v) in list(stats.items())[:10]:
            want_m = W0[name + "/moving_mean"] * 0.997 + m * 0.003
            want_v = W0[name + "/moving_variance"] * 0.997 + v * 0.003
            assert rel_err(W1[name + "/moving_mean"], want_m) < 1e-4, name
            assert rel_err(W1[name + "/moving_variance"], want_v) < 1e-4, name
    # backbone untouched
    assert np.array_equal(W1["stem_conv/kernel"], W0["stem_conv/kernel"])
    # compact device targets give the same step
    model2 = efficientdet(phi, num_classes=C, weighted_bifpn=weighted, freeze_bn=freeze_bn,
                          image_size=size, dtype="fp32", drop_connect_rate=0, just_training_model=True)
    model2.set_weights_dict(W0)
    for i in range(1, EFFICIENTNET_DEPTHS[phi]):
        model2.layers[i].trainable = False
    model2.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9))
    from efficientdet_b200.utils.anchors import anchor_targets_device
    r, _, st, cl = anchor_targets_device(anchors, [(size, size, 3)] * B, ann, C, dense_labels=False,
                                         compact=True)
    t2 = model2.train_on_batch(img, (r, st, cl))
    assert abs(t2[0] - total) / total < 1e-6
    k = "class_head/pyramid_classification/kernel"
    assert np.array_equal(model2.net.grads[k].cpu().numpy(), net.grads[k].cpu().numpy())


def test_training_requires_frozen_backbone():
    from efficientdet_b200.model import efficientdet
    model = efficientdet(0, num_classes=3, image_size=128, just_training_model=True)
    model.compile()
    with pytest.raises(NotImplementedError):
        model.train_on_batch(np.zeros((1, 128, 128, 3), np.float32),
                             [np.zeros((1, 3069, 5), np.float32), np.zeros((1, 3069, 4), np.float32)])
