"""The CUDA path against fixtures produced by EXECUTING the reference's own model.py / efficientnet.py /
layers.py / utils/tpu.py (tests/golden/make_golden_graph.py, make_golden_losses.py): product vs reference
wiring directly, no oracle in between.  Tolerances (BASELINE.json north_star): 1e-4 (fp32) / 2e-2 (bf16),
relative = max|got - want| / max|want| per tensor; an RMS-normalised bound (||got-want|| / ||want||) is
asserted next to it."""
import numpy as np
import pytest
import torch

from test_oracle_graph_golden import CASES, load_case
from util_model import rel_err, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", CASES)
def test_forward_matches_reference_graph(tag, dtype):
    from efficientdet_b200.model import efficientdet
    z, phi, C, weighted, S, W, img = load_case(tag)
    model = efficientdet(phi, num_classes=C, weighted_bifpn=weighted, image_size=S, dtype=dtype,
                         just_training_model=True)      # default drop_connect_rate=0.2: identity at inference
    mine = model.get_weights_dict()
    # the weight manifest of the reference graph == the product's state dict (names and Keras shapes)
    assert set(mine) == set(W), (sorted(set(mine) ^ set(W))[:10])
    for k in W:
        assert tuple(mine[k].shape) == tuple(W[k].shape), k
    model.set_weights_dict(W, strict=True)
    plan = model.net.plan(2, keep_taps=True)
    reg, cls = plan.forward(torch.from_numpy(img).cuda())
    torch.cuda.synchronize()
    tol, tol2 = (1e-4, 1e-4) if dtype == "fp32" else (2e-2, 1e-2)
    bad = {}
    for k in z.files:
        if k.startswith(("C", "BiFPN_")):
            got = plan.tensor(plan.taps[k]).float().cpu().numpy()
        elif k == "regression":
            got = reg.cpu().numpy()
        elif k == "classification":
            got = cls.cpu().numpy()
        else:
            continue
        e, e2 = rel_err(got, z[k]), rel_l2(got, z[k])
        if not (e < tol and e2 < tol2):
            bad[k] = (e, e2)
    assert not bad, bad


def test_losses_match_reference_code():
    """effdet_detection_losses vs the loss values / gradients computed by the reference's utils/tpu.py."""
    import os
    from efficientdet_b200 import _lib
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses.npz"))
    lab, p, reg_t, reg_p = z["labels"], z["pred"], z["reg_t"], z["reg_p"]
    B, N, C = p.shape
    lib = _lib.load()
    d = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)
    state = lab[..., -1]
    clsid = np.where(state == 1, lab[..., :-1].argmax(-1), -1)
    pd, rd, rtd, labd = d(p), d(reg_p), d(reg_t), d(lab)
    st, cl = d(state, torch.int8), d(clsid, torch.int32)
    dcls = torch.empty((B, N, C), device="cuda")
    dreg = torch.empty((B, N, 4), device="cuda")
    out8 = torch.zeros(8, device="cuda")
    wsb = lib.effdet_detection_losses_workspace_size()
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    p32 = p.astype(np.float32).astype(np.float64)
    for ftag, stag in (("a", "a"), ("b", "b")):
        alpha, gamma = [float(v) for v in z["focal_%s_params" % ftag]]
        lam = float(z["sl1_%s_lambda" % stag])
        for dense in (True, False):
            _lib.call("effdet_detection_losses", pd.data_ptr(), rd.data_ptr(), rtd.data_ptr(),
                      labd.data_ptr() if dense else None, st.data_ptr(), cl.data_ptr(), B, N, C, alpha, gamma,
                      lam, 1.0, dcls.data_ptr(), dreg.data_ptr(), out8.data_ptr(), ws.data_ptr(), wsb,
                      None, None, None, 0, 0, 0, _lib.stream_ptr())
            o = out8.cpu().numpy()
            assert abs(o[0] - float(z["focal_%s" % ftag])) < 1e-4 * float(z["focal_%s" % ftag])
            assert abs(o[1] - float(z["sl1_%s" % stag])) < 1e-4 * float(z["sl1_%s" % stag])
            # the kernel returns d/d(logit); the reference gradient is d/dp: chain through the sigmoid
            want = z["focal_%s_grad" % ftag] * p32 * (1 - p32)
            g = dcls.cpu().numpy()
            skip = np.zeros_like(g, bool)
            skip[0, :5] = True                  # clip-boundary / p == .5 points: sub-gradient conventions
            assert rel_err(g[~skip], want[~skip]) < 1e-4
            assert rel_err(dreg.cpu().numpy(), z["sl1_%s_grad" % stag]) < 1e-5


def test_wbifpn_add_matches_reference_layer():
    import os
    from efficientdet_b200.layers import wBiFPNAdd
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses.npz"))
    for n in (2, 3):
        layer = wBiFPNAdd(name="w_bi_fpn_add")
        xs = [x.astype(np.float32) for x in z["fuse%d_x" % n]]
        layer.build([x.shape for x in xs])
        layer.set_weights([z["fuse%d_w" % n].astype(np.float32)])
        y = layer.call(xs)
        assert rel_err(np.asarray(y), z["fuse%d_y" % n]) < 1e-6
