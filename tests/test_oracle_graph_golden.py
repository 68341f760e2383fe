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


def load_case(tag):
    z = np.load(os.path.join(HERE, "golden", "graph_%s.npz" % tag))
    phi, C, weighted, S, seed = [int(v) for v in z["meta"]]
    names = [str(n) for n in z["weight_names"]]
    shapes = [tuple(int(s) for s in str(t).split(",")) for t in z["weight_shapes"]]
    W = {n: golden_weight(n, s, seed) for n, s in zip(names, shapes)}
    img = np.random.default_rng(1000 + seed).standard_normal((2, S, S, 3)).astype(np.float32)
    return z, phi, C, bool(weighted), S, W, img


@pytest.mark.parametrize("tag", CASES)
def test_oracle_forward_matches_reference_graph(tag):
    from oracle import graph
    z, phi, C, weighted, S, W, img = load_case(tag)
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
