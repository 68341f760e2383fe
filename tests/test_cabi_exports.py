"""CPU: the C-ABI library loads (no GPU needed) and exports every symbol the header declares."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "effdet_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(effdet_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from efficientdet_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert lib.effdet_version() >= 100


def test_python_binding_covers_header():
    from efficientdet_b200 import _lib
    _lib.load()
    bound = set(_lib._SIGNATURES) | {"effdet_last_error", "effdet_version", "effdet_launch_count",
                                     "effdet_filter_detections_workspace_size",
                                     "effdet_detection_losses_workspace_size",
                                     "effdet_colreduce_blocks", "effdet_dw_wgrad_blocks",
                                     "effdet_conv_wgrad_splits", "effdet_conv_wgrad_tc_splits", "effdet_conv_wgrad_tc_fuses_bias", "effdet_dwconv_se_blocks",
                                     "effdet_conv_tc_block_n", "effdet_conv_weight_panel_elems", "effdet_conv_weight_panel_split_elems",
                                     "effdet_se_backward_blocks", "effdet_se_bn_backward_blocks", "effdet_dw_backward_blocks",
                                     "effdet_stem_wgrad_blocks", "effdet_plan_num_weights",
                                     "effdet_plan_num_anchors", "effdet_plan_num_launches", "effdet_plan_dry_run",
                                     "effdet_replay_num_launches"}
    missing = [n for n in _declared() if n not in bound]
    assert not missing, missing


def test_host_anchor_table_bit_exact(golden):
    from efficientdet_b200.utils import anchors as A
    for shp in ((128, 128), (96, 160), (100, 150)):
        assert np.array_equal(A.anchors_for_shape(shp), golden["anchors_%dx%d" % shp])
    for s in (16, 32, 64, 128, 256, 512, 48):
        assert np.array_equal(A.generate_anchors(s), golden["gen_%d" % s])
    import hashlib
    for i, S in enumerate(golden["model_sizes"]):
        a = A.anchors_for_shape((int(S), int(S)))
        assert hashlib.sha256(a.tobytes()).hexdigest() == str(golden["model_sha_f64"][i])


def test_invalid_arguments_raise_without_gpu():
    from efficientdet_b200 import _lib
    import pytest
    out = np.zeros((1, 4))
    hw = np.array([[4, 4]], np.int32); one = np.array([32], np.int32)
    r = np.array([1.0]); s = np.array([1.0])
    with pytest.raises(_lib.CapacityError):
        _lib.call("effdet_anchors_for_shape_host", hw.ctypes.data, one.ctypes.data, one.ctypes.data,
                  1, r.ctypes.data, 1, s.ctypes.data, 1, out.ctypes.data, 1)
    with pytest.raises(ValueError):
        _lib.call("effdet_anchors_for_shape_host", None, one.ctypes.data, one.ctypes.data,
                  1, r.ctypes.data, 1, s.ctypes.data, 1, out.ctypes.data, 1)


def test_replay_thunks_are_up_to_date():
    """csrc/replay_thunks.inc (dispatch table of the compiled-plan replayer) is generated from the ctypes signature
    table; the committed file must be what the generator produces now."""
    from efficientdet_b200 import _gen_replay_thunks as g
    assert open(g.OUT).read() == g.generate()
    assert "thunk_effdet_conv2d" in g.generate() and "thunk_effdet_sgd_momentum_step_dev_lr" in g.generate()
