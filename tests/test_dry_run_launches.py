"""Host side of EVERY launch, on the CPU: the launch lists of the inference and training plans are built with their
real arguments (host addresses standing in for device addresses) and each launch is issued through the real C entry
point.  Without a CUDA driver the library cannot launch anything -- it has no CPU fallback -- so every call must end
in EFFDET_E_CUDA at its first CUDA runtime call ("driver version is insufficient"); what runs BEFORE that point is the
entry point's own argument validation and launch configuration: tile shapes, tensor-map construction
(EFFDET_DRY_RUN swaps cuTensorMapEncodeTiled, which needs a driver, for a checker of its documented argument rules,
csrc/tma.cuh), shared-memory budgets, split counts, channel-vector limits.  Any other outcome (EFFDET_E_INVALID /
_UNSUPPORTED / _CAPACITY) is a shape the lowering produces and the library rejects -- on a GPU it would be a failed
step.  Swept over D0..D6 at the reference's image sizes, odd class counts, batch sizes 1..128, bf16 / fp32,
weighted / plain BiFPN, frozen / trained backbone.

This sweep is how the 6- and 7-class failure of the D0 class head (230 KiB of shared memory in the halo form of the
tensor-core convolution, found on a B200 by examples/detect_host.c) is kept from coming back: with `halo_fits()` disabled the
sweep reports "235608 bytes of shared memory exceed a CTA's 227 KiB" for `class_head/pyramid_classification`."""
import os

import pytest
import torch

IMAGE_SIZES = (512, 640, 768, 896, 1024, 1280, 1408)          # model.py:29


@pytest.fixture(scope="module")
def dry(request):
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the launches would run")
    from efficientdet_b200 import _lib, engine
    saved = (_lib.stream_ptr, _lib.call, engine.Plan.__init__, os.environ.get("EFFDET_DRY_RUN"))

    def structure_only(self, net, batch, reuse_buffers=True, keep_taps=False, u8_input=False):
        self.net, self.u8_input, self.B, self.dev, self.dtype = net, bool(u8_input), int(batch), net.device, net.dtype
        self.ops, self.vals, self.taps, self.keep_taps = [], [], {}, keep_taps
        self.reuse = reuse_buffers and not keep_taps
        self._keepalive, self.graph = [], None
        self._build()
        self._assign_buffers()
        for op in self.ops:
            op.fn = op.make()

    os.environ["EFFDET_DRY_RUN"] = "1"
    _lib.stream_ptr = lambda device=None: 0
    _lib.call = lambda *a, **k: 0          # weight-preparation launches issued while BUILDING a network / plan
    engine.Plan.__init__ = structure_only
    try:
        yield _lib
    finally:
        _lib.stream_ptr, _lib.call, engine.Plan.__init__ = saved[:3]
        if saved[3] is None:
            os.environ.pop("EFFDET_DRY_RUN", None)
        else:
            os.environ["EFFDET_DRY_RUN"] = saved[3]


def _issue_all(_lib, plan, where):
    n = 0
    for op in plan.ops:
        with pytest.raises(_lib.EffdetError) as e:
            op.fn(0)
        msg = str(e.value)
        assert e.value.code == _lib.E_CUDA and "driver version is insufficient" in msg, (
            "%s: launch %s (%s) was rejected before reaching CUDA: %s" % (where, op.name, op.kind, msg))
        n += 1
    return n


def _plans(phi, size, batch, classes, dtype, weighted, which):
    from efficientdet_b200 import engine, train
    from efficientdet_b200.model import efficientdet
    m = efficientdet(phi, num_classes=classes, image_size=size, weighted_bifpn=weighted, just_training_model=True,
                     device="cpu", dtype=dtype)
    if "i" in which:
        yield "inference", engine.Plan(m.net, batch)
    if "u" in which:
        yield "inference uint8", engine.Plan(m.net, batch, u8_input=True)
    if "t" in which:
        yield "training", train.TrainPlan(m.net, batch, train_backbone=True)
    if "f" in which:
        yield "training frozen", train.TrainPlan(m.net, batch, train_backbone=False, u8_input=True)


def test_error_type_exposes_the_code(dry):
    assert issubclass(dry.EffdetError, Exception) and dry.E_CUDA == -2


@pytest.mark.parametrize("phi", range(7))
def test_reference_image_sizes(dry, phi):
    """efficientdet(phi) at image_sizes[phi], 90 classes, batch 1: configs 1, 3, 4, 5 of BASELINE.json."""
    total = 0
    for dtype in ("bf16", "fp32"):
        for name, p in _plans(phi, IMAGE_SIZES[phi], 1, 90, dtype, phi >= 3, "iutf"):
            total += _issue_all(dry, p, "D%d %s %s" % (phi, dtype, name))
    assert total > 1000


@pytest.mark.parametrize("phi", [0, 1, 3, 4, 6])
def test_class_counts(dry, phi):
    """The class-head final convolution has 9 * classes output channels (model.py:330-345): every N-tile width, TMA
    and non-TMA output strides, focal-loss vector widths, tail segment counts."""
    full = (1, 2, 3, 5, 6, 7, 9, 11, 20, 28, 29, 57, 80, 90, 91, 113, 200, 226, 452)
    for classes in (full if phi == 0 else (1, 6, 7, 20, 91, 452)):
        for dtype in ("bf16", "fp32"):
            for name, p in _plans(phi, 128, 2, classes, dtype, False, "if"):
                _issue_all(dry, p, "D%d %d classes %s %s" % (phi, classes, dtype, name))


def test_class_count_limit_is_reported_when_the_plan_is_built(dry):
    """Training reduces the class-head bias gradient with one column vector per thread (<= 1024 per block; odd
    9 * classes rows are folded 2 or 4 times to reach the vector width): larger heads are rejected by the lowering,
    not at the first step.  bf16 training on a BiFPN of width <= 64 takes that gradient from the tensor-core
    weight-gradient launch and inference has no such reduction: no limit there."""
    for classes in (115, 230, 456, 601):
        with pytest.raises(ValueError, match="num_classes"):
            for _ in _plans(0, 128, 2, classes, "fp32", False, "f"):
                pass
    for name, p in _plans(0, 128, 2, 601, "bf16", False, "if"):
        _issue_all(dry, p, "601 classes " + name)
    with pytest.raises(ValueError, match="num_classes"):
        for _ in _plans(1, 128, 2, 601, "bf16", False, "f"):          # D1: W = 88
            pass


@pytest.mark.parametrize("phi", [0, 2, 4])
def test_batch_sizes(dry, phi):
    for batch in (1, 2, 3, 5, 8, 16, 33, 64, 128):
        for name, p in _plans(phi, 256, batch, 20, "bf16", True, "it"):
            _issue_all(dry, p, "D%d batch %d %s" % (phi, batch, name))


def test_other_image_sizes(dry):
    """Any multiple of 128 (pyramid levels with odd extents: 384 -> 48, 24, 12, 6, 3)."""
    for size in (128, 384, 1152):
        for name, p in _plans(1, size, 2, 20, "bf16", True, "it"):
            _issue_all(dry, p, "D1 %d px %s" % (size, name))


# ---------------------------------------------------------------- the C++ lowering (csrc/plan.cu)
def _c_plan(lib, phi, size, batch, classes, weighted, dtype, u8):
    import ctypes
    plan = ctypes.c_void_p()
    rc = lib.effdet_plan_create(phi, size, batch, classes, int(weighted), 1 if dtype == "bf16" else 0,
                                1 if u8 else 0, ctypes.byref(plan))
    assert rc == 0, lib.effdet_last_error()
    return plan


@pytest.mark.parametrize("phi", range(7))
def test_cpp_plan_matches_python_lowering_and_passes_dry_run(dry, phi):
    """effdet_plan_create (the plan-level C ABI, lowered in C++) against the Python lowering for every model size:
    the same weight manifest (Keras names, shapes, order -- what load_weights(by_name=True) matches), the same number
    of launches per forward, and every one of its launches passes its entry point's host side (effdet_plan_dry_run).
    The GPU test (tests/test_gpu_plan_cabi.py) compares outputs bit for bit on four configurations; this covers the
    structure of all of them."""
    import ctypes
    from efficientdet_b200 import engine
    from efficientdet_b200.model import efficientdet
    lib = dry.load()
    for weighted, classes, size in ((False, 90, IMAGE_SIZES[phi]), (True, 6, 256)):
        for dtype in ("bf16", "fp32"):
            m = efficientdet(phi, num_classes=classes, image_size=size, weighted_bifpn=weighted,
                             just_training_model=True, device="cpu", dtype=dtype)
            mine = m.get_weights_dict()
            for u8 in (False, True):
                plan = _c_plan(lib, phi, size, 1, classes, weighted, dtype, u8)
                try:
                    manifest = []
                    for i in range(lib.effdet_plan_num_weights(plan)):
                        name, nd, dims = ctypes.c_char_p(), ctypes.c_int(), (ctypes.c_int * 4)()
                        assert lib.effdet_plan_weight_info(plan, i, ctypes.byref(name), ctypes.byref(nd), dims) == 0
                        manifest.append((name.value.decode(), tuple(dims[:nd.value])))
                    want = [(k, tuple(v.shape)) for k, v in mine.items() if not k.startswith("boxes/")]
                    assert manifest == want
                    py = engine.Plan(m.net, 1, u8_input=u8)
                    assert lib.effdet_plan_num_launches(plan) == len(py.ops)
                    assert lib.effdet_plan_num_anchors(plan) == py.N
                    n = ctypes.c_int()
                    rc = lib.effdet_plan_dry_run(plan, ctypes.byref(n))
                    assert rc == 0, "D%d %s: %s" % (phi, dtype, lib.effdet_last_error().decode())
                    assert n.value > lib.effdet_plan_num_launches(plan)        # + BatchNorm folds and weight panels
                finally:
                    lib.effdet_plan_destroy(plan)


def test_cpp_plan_dry_run_class_counts_and_batches(dry):
    import ctypes
    lib = dry.load()
    for phi, size in ((0, 256), (1, 128), (4, 128)):
        for classes in (1, 2, 3, 5, 6, 7, 9, 20, 91, 200, 601):
            for batch in (1, 3, 64):
                plan = _c_plan(lib, phi, size, batch, classes, True, "bf16", True)
                rc = lib.effdet_plan_dry_run(plan, None)
                msg = lib.effdet_last_error().decode()
                lib.effdet_plan_destroy(plan)
                assert rc == 0, "D%d %d classes batch %d: %s" % (phi, classes, batch, msg)
