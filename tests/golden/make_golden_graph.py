"""Generates tests/golden/graph_*.npz by EXECUTING THE REFERENCE's own graph-construction code:
/root/reference/model.py (efficientdet, build_BiFPN / build_wBiFPN, build_regress_head / build_class_head),
efficientnet.py (EfficientNet, mb_conv_block), tfkeras.py, layers.py (wBiFPNAdd), initializers.py are imported
UNMODIFIED; `tensorflow` / `tensorflow.keras` / `keras_applications` are replaced by tests/golden/keras_stub.py
(torch-CPU float64 layers, TF op semantics restated there once).  This pins the oracle's -- and the product's --
topology, layer names, widths, skip / drop conditions, fusion-input order and head order on the reference
itself; the arithmetic inside TensorFlow's kernels stays restated.

Only runs in the build container (needs /root/reference).  Run:  python tests/golden/make_golden_graph.py

Each fixture holds: the weight manifest (every "<layer>/<weight>" the reference created, with its shape, in
creation order), the number of Keras layers the backbone call created (+ the input layer) -- train_tpu.py:24's
EFFICIENTNET_DEPTHS -- and, for the image batch default_rng(1000 + seed).standard_normal((2,S,S,3)) with weights = tests/util_model.golden_weight,
the backbone features C3..C5, every BiFPN layer's five outputs, regression and classification (float32).
Input size: the reference declares Input((image_sizes[phi],)*2 + (3,)) (model.py:368-371); the convolutional
graph is size-agnostic and the stub evaluates it on 128x128 images to keep the fixtures small.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import keras_stub as ks  # noqa: E402
from util_model import golden_weight  # noqa: E402

ks.install()
sys.path.insert(0, "/root/reference")
import model as ref_model  # noqa: E402  (the reference's model.py, unmodified)

CASES = [  # (tag, phi, num_classes, weighted, image side, seed)
    ("d0", 0, 20, False, 128, 11),
    ("d0w", 0, 8, True, 128, 12),
    ("d1", 1, 4, False, 128, 13),       # B1: block1b = skip block without expansion
    ("d3w", 3, 6, True, 128, 14),       # head depth 4, 5 BiFPN layers
]


def run_case(tag, phi, C, weighted, S, seed, out_dir=HERE):
    ks.reset()
    ks.WEIGHTS = lambda key, shape: golden_weight(key, shape, seed)
    rec = {}
    # pass-through recorders around the reference's own builders (the table / module attributes are
    # wrapped, the builder code itself runs as written)
    orig_bb = ref_model.backbones[phi]
    orig_b, orig_w = ref_model.build_BiFPN, ref_model.build_wBiFPN

    def bb(*a, **k):
        n0 = len(ks.CREATED)
        feats = orig_bb(*a, **k)
        rec["n_backbone_layers"] = len(ks.CREATED) - n0 + 1       # + the input layer
        rec["feats"] = list(feats)
        return feats

    def wrap(fn):
        def g(features, num_channels, idx, **k):
            outs = fn(features, num_channels, idx, **k)
            rec.setdefault("bifpn", []).append(list(outs))
            return outs
        return g
    ref_model.backbones[phi] = bb
    ref_model.build_BiFPN, ref_model.build_wBiFPN = wrap(orig_b), wrap(orig_w)
    try:
        m = ref_model.efficientdet(phi, num_classes=C, weighted_bifpn=weighted, just_training_model=True)
    finally:
        ref_model.backbones[phi] = orig_bb
        ref_model.build_BiFPN, ref_model.build_wBiFPN = orig_b, orig_w
    rng = np.random.default_rng(1000 + seed)
    img = rng.standard_normal((2, S, S, 3)).astype(np.float32)
    vals = {}
    reg, cls = m.predict(img, record=vals)
    out = {"regression": reg.numpy().astype(np.float32),
           "classification": cls.numpy().astype(np.float32)}
    for j, t in enumerate(rec["feats"]):
        if j < 2:
            continue                      # C1 / C2 feed nothing (model.py:95,201) and are the largest tensors
        out["C%d" % (j + 1)] = vals[id(t)].numpy().astype(np.float32)
    assert len(rec["bifpn"]) == 2 + phi
    for i, outs in enumerate(rec["bifpn"]):
        for j, t in enumerate(outs):
            out["BiFPN_%d_P%d" % (i, j + 3)] = vals[id(t)].numpy().astype(np.float32)
    W = m.all_weights()
    names = list(W.keys())
    out["weight_names"] = np.array(names)
    out["weight_shapes"] = np.array([",".join(str(s) for s in W[k].shape) for k in names])
    out["n_backbone_layers"] = np.int64(rec["n_backbone_layers"])
    out["layer_names"] = np.array([l.name for l in ks.CREATED])
    out["meta"] = np.array([phi, C, int(weighted), S, seed], np.int64)
    # known answer of the reference's PriorProbability initializer (initializers.py:24)
    init = ks.INITIALIZERS["class_head/pyramid_classification/bias"]
    out["prior_bias"] = np.float64(float(init((1,))[0]))
    assert np.isfinite(out["regression"]).all() and np.isfinite(out["classification"]).all()
    np.savez_compressed(os.path.join(out_dir, "graph_%s.npz" % tag), **out)
    print(tag, "layers(backbone)=%d" % rec["n_backbone_layers"], "weights=%d" % len(names),
          "reg", out["regression"].shape, float(np.abs(out["regression"]).max()),
          "cls", out["classification"].shape, float(out["classification"].min()),
          float(out["classification"].max()),
          "P3", float(np.abs(out["BiFPN_%d_P3" % (1 + phi)]).max()))


if __name__ == "__main__":
    # no arguments: regenerate the committed fixtures.  `--out DIR tag:phi:classes:weighted:side:seed ...` writes
    # further cases elsewhere (tests/test_oracle_graph_golden.py runs the remaining model sizes live this way)
    if len(sys.argv) > 1 and sys.argv[1] == "--out":
        for spec in sys.argv[3:]:
            tag, rest = spec.split(":", 1)
            run_case(tag, *[int(v) for v in rest.split(":")], out_dir=sys.argv[2])
    else:
        for c in CASES:
            run_case(*c)
