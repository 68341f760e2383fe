"""Pins oracle/graph.py (and the oracle's structure tables) on fixtures produced by EXECUTING the reference's
own graph-construction code (tests/golden/make_golden_graph.py: /root/reference/model.py + efficientnet.py +
layers.py unmodified under a torch-backed Keras stand-in).  CPU only."""
import os

import numpy as np
import pytest
import torch

from util_model import golden_weight

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["d0", "d0w", "d1", "d3w"]
EFFICIENTNET_DEPTHS = [227, 329, 329, 374, 464, 566, 656]      # train_tpu.py:24


def load_case(tag, directory=None):
    z = np.load(os.path.join(directory or os.path.join(HERE, "golden"), "graph_%s.npz" % tag))
    phi, C, weighted, S, seed = [int(v) for v in z["meta"]]
    names = [str(n) for n in z["weight_names"]]
    shapes = [tuple(int(s) for s in str(t).split(",")) for t in z["weight_shapes"]]
    W = {n: golden_weight(n, s, seed) for n, s in zip(names, shapes)}
    img = np.random.default_rng(1000 + seed).standard_normal((2, S, S, 3)).astype(np.float32)
    return z, phi, C, bool(weighted), S, W, img


def _check_forward(tag, directory=None):
    from oracle import graph
    z, phi, C, weighted, S, W, img = load_case(tag, directory)
    taps = {}
    with torch.no_grad():
        reg, cls = graph.forward(W, img, phi, C, weighted, dtype=torch.float64, taps=taps)
    got = dict(taps, regression=reg, classification=cls)
    checked = 0
    for k in z.files:
        if k.startswith(("C", "BiFPN_")) or k in ("regression", "classification"):
            want = z[k].astype(np.float64)
            g = got[k].numpy()
            assert g.shape == want.shape, (k, g.shape, want.shape)
            err = np.abs(g - want).max() / max(np.abs(want).max(), 1e-30)
            assert err < 2e-6, (tag, k, err)         # fixtures are stored as float32
            checked += 1
    assert checked == 3 + 5 * (2 + phi) + 2
    assert int(z["n_backbone_layers"]) == EFFICIENTNET_DEPTHS[phi] == graph.keras_layer_count(phi)


@pytest.mark.parametrize("tag", CASES)
def test_oracle_forward_matches_reference_graph(tag):
    _check_forward(tag)


@pytest.mark.skipif(not os.path.exists("/root/reference/model.py"),
                    reason="build container only: executes the reference's model.py from where it lies")
def test_remaining_model_sizes_against_the_reference_executed_live(tmp_path, monkeypatch):
    """The committed fixtures cover D0, D0 weighted, D1, D3 weighted; here the reference's own model.py builds and
    runs D2 weighted, D4, D5 weighted and D6 (tests/golden/make_golden_graph.py in a subprocess, fixtures in a
    temporary directory) and the oracle is compared the same way: every pyramid level of every BiFPN layer, both
    heads, the backbone's Keras layer count -- so the wiring of ALL seven model sizes is pinned on executed
    reference code."""
    import subprocess
    import sys
    cases = ["d2w:2:5:1:128:21", "d4:4:3:0:128:22", "d5w:5:4:1:128:23", "d6:6:2:0:128:24"]
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_golden_graph.py"), "--out", str(tmp_path)]
                       + cases, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:]
    for c in cases:
        _check_forward(c.split(":")[0], str(tmp_path))
    # the PRODUCT's builder against the same executed reference: the set of weight names and their shapes (what
    # load_weights(by_name=True) matches); built on the CPU with the launches stubbed
    from efficientdet_b200 import _lib
    from efficientdet_b200.model import efficientdet
    monkeypatch.setattr(_lib, "stream_ptr", lambda device=None: 0)
    monkeypatch.setattr(_lib, "call", lambda *a, **k: 0)
    for c in cases:
        z, phi, C, weighted, S, W, img = load_case(c.split(":")[0], str(tmp_path))
        m = efficientdet(phi, num_classes=C, weighted_bifpn=weighted, image_size=S, just_training_model=True,
                         device="cpu", dtype="fp32")
        mine = {k: tuple(v.shape) for k, v in m.get_weights_dict().items() if not k.startswith("boxes/")}
        assert mine == {n: tuple(W[n].shape) for n in (str(x) for x in z["weight_names"])}, c
        assert m.backbone_depth == EFFICIENTNET_DEPTHS[phi]


@pytest.mark.parametrize("tag", CASES)
def test_structure_matches_reference_graph(tag):
    """Layer count of the backbone (train_tpu.py:24) and the weight manifest (names, shapes) the reference
    created, against the oracle's block table / layer-count model."""
    from oracle import graph
    z, phi, C, weighted, S, W, img = load_case(tag)
    assert int(z["n_backbone_layers"]) == EFFICIENTNET_DEPTHS[phi] == graph.keras_layer_count(phi)
    blocks, taps = graph.block_list(phi)
    for b in blocks:
        p = b["prefix"]
        cm = b["cin"] * b["expand"]
        assert (p + "expand_conv/kernel" in W) == (b["expand"] != 1)
        assert W[p + "dwconv/depthwise_kernel"].shape == (b["k"], b["k"], cm, 1)
        assert W[p + "se_reduce/kernel"].shape == (1, 1, cm, b["se"])
        assert W[p + "project_conv/kernel"].shape == (1, 1, cm, b["cout"])
        names = set(str(n) for n in z["layer_names"])
        assert (p + "add" in names) == b["skip"]
        assert (p + "drop" in names) == (b["skip"] and b["num"] > 0)
    n_fuse = sum(1 for k in W if k.startswith("w_bi_fpn_add"))
    assert n_fuse == (8 * (2 + phi) if weighted else 0)
    depth = 3 + phi // 3
    assert ("box_head/regress_head_conv_%d/kernel" % (depth - 1)) in W
    assert ("box_head/regress_head_conv_%d/kernel" % depth) not in W
    assert W["class_head/pyramid_classification/kernel"].shape == (3, 3, graph.W_BIFPNS[phi], 9 * C)
    # PriorProbability(0.01) evaluated by the reference's own initializer class (initializers.py:24)
    assert abs(float(z["prior_bias"]) - (-np.log(99.0))) < 1e-12
