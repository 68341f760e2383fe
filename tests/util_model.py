"""Shared helpers for the GPU parity tests (synthetic weights per SURVEY.md section 8(d))."""
import numpy as np


def perturb_weights(model, seed=2024):
    """Exercise BN / fusion paths: gamma~U(.5,1.5), beta~N(0,.1), moving_mean~N(0,.1),
    moving_variance~U(.5,1.5), fusion w~U(-.2,1), biases~N(0,.05)."""
    rng = np.random.default_rng(seed)
    d = model.get_weights_dict()
    for k, v in d.items():
        if k.endswith("/gamma"):
            d[k] = rng.uniform(0.5, 1.5, v.shape).astype(np.float32)
        elif k.endswith("/beta") or k.endswith("/moving_mean"):
            d[k] = rng.normal(0, 0.1, v.shape).astype(np.float32)
        elif k.endswith("/moving_variance"):
            d[k] = rng.uniform(0.5, 1.5, v.shape).astype(np.float32)
        elif k.startswith("w_bi_fpn_add"):
            d[k] = rng.uniform(-0.2, 1.0, v.shape).astype(np.float32)
        elif k.endswith("/bias") and "pyramid_classification" not in k:
            d[k] = rng.normal(0, 0.05, v.shape).astype(np.float32)
        elif k.startswith(("box_head", "class_head")) and k.endswith("/kernel"):
            # N(0,.01) heads give ~1e-5 outputs; widen so that errors are visible
            d[k] = (rng.standard_normal(v.shape) * np.sqrt(2.0 / (9 * v.shape[2]))).astype(np.float32)
    model.set_weights_dict(d)
    return d


def rel_err(got, want):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def rel_l2(got, want):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30))


def golden_weight(key, shape, seed=0):
    """Deterministic synthetic value of the Keras weight `key` ("[model/]layer/weight"): a pure function of
    (key, shape, seed), shared by tests/golden/make_golden_graph.py (which feeds it to the reference's own
    graph-construction code) and the parity tests (which feed it to the oracle and to the CUDA path), so the
    fixtures only have to carry the weight manifest (names + shapes) and the outputs.  Scales keep activations
    O(1) through ~100 layers: N(0, 1/fan_in) kernels, BN gamma~U(.5,1.5), beta / moving_mean~N(0,.1),
    moving_variance~U(.5,1.5), biases~N(0,.05), fusion weights~U(-.2,1) (some negative: exercises the relu)."""
    import zlib
    rng = np.random.default_rng([zlib.crc32(key.encode()), int(seed)])
    leaf = key.rsplit("/", 1)[1]
    shape = tuple(int(s) for s in shape)
    if leaf == "kernel":
        fan_in = shape[0] * shape[1] * shape[2]
        v = rng.standard_normal(shape) * np.sqrt(1.0 / fan_in)
    elif leaf == "depthwise_kernel":
        v = rng.standard_normal(shape) * np.sqrt(1.0 / (shape[0] * shape[1]))
    elif leaf == "gamma":
        v = rng.uniform(0.5, 1.5, shape)
    elif leaf in ("beta", "moving_mean"):
        v = rng.normal(0, 0.1, shape)
    elif leaf == "moving_variance":
        v = rng.uniform(0.5, 1.5, shape)
    elif leaf == "bias":
        v = rng.normal(0, 0.05, shape) - (2.0 if "pyramid_classification" in key else 0.0)
    elif leaf.startswith("w_bi_fpn_add"):
        v = rng.uniform(-0.2, 1.0, shape)
    else:
        raise KeyError(key)
    return v.astype(np.float32)
