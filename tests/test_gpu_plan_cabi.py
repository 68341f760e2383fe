"""The plan-level C ABI (effdet_plan_create / bind_weights / forward / detect, csrc/plan.cu) driven purely through
ctypes with numpy HOST buffers -- the way a non-Python host would use it -- against the Python-side plan:
bit-identical detections (both issue the same launches of the same library)."""
import ctypes

import numpy as np
import pytest
import torch

from util_model import golden_weight

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("phi,size,B,C,weighted,dtype,u8", [
    (0, 256, 2, 6, False, "fp32", False), (0, 512, 3, 20, True, "bf16", False), (1, 256, 2, 4, True, "bf16", True),
    (3, 256, 1, 8, False, "bf16", False)])
def test_detect_through_the_c_abi_matches_predict_on_batch(phi, size, B, C, weighted, dtype, u8):
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.plan import CPlan
    from efficientdet_b200.utils.anchors import anchors_for_shape
    plan = CPlan(phi, size, B, C, weighted, dtype, u8_input=u8)
    manifest = plan.weight_manifest()
    anchors = anchors_for_shape((size, size))
    assert plan.N == anchors.shape[0]
    model, pmodel = efficientdet(phi, num_classes=C, weighted_bifpn=weighted, image_size=size, dtype=dtype,
                                 score_threshold=0.3, anchors=anchors)
    mine = model.get_weights_dict()
    # the C++ manifest == the Python state dict == (tests/test_gpu_golden_graph.py) the reference graph's weights
    assert [n for n, _ in manifest] == [k for k in mine if not k.startswith("boxes/")]
    assert all(tuple(mine[n].shape) == s for n, s in manifest)
    W = {n: golden_weight(n, s, 9) for n, s in manifest}
    model.set_weights_dict(W, strict=True)
    plan.bind_weights_host(W)
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (B, size, size, 3), dtype=np.uint8) if u8 else \
        rng.standard_normal((B, size, size, 3)).astype(np.float32)
    r0, c0 = model.predict_on_batch(img)
    thr = float(np.quantile(c0, 0.999))                  # ~0.1 % of the (anchor, class) scores are candidates
    pmodel.score_threshold = thr
    want = pmodel.predict_on_batch([img])
    got = plan.detect_host(img, score_threshold=thr)
    assert (want[2] >= 0).sum() > 10                     # the comparison is not vacuous
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    # anchors as an input (inference.py:57-59) and a second call on the captured graph
    got2 = plan.detect_host(img, anchors=anchors[None].astype(np.float32), score_threshold=thr)
    for g, w in zip(got2, want):
        assert np.array_equal(g, w)
    # raw head outputs through effdet_forward with device pointers
    reg = torch.empty((B, plan.N, 4), device="cuda")
    cls = torch.empty((B, plan.N, C), device="cuda")
    x = torch.from_numpy(img).cuda()
    plan.forward_device(x.data_ptr(), reg.data_ptr(), cls.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(reg.cpu().numpy(), r0) and np.array_equal(cls.cpu().numpy(), c0)
    plan.close()


def test_plan_rejects_unbound_weights_and_bad_arguments():
    from efficientdet_b200 import _lib
    from efficientdet_b200.plan import CPlan
    with pytest.raises(ValueError):
        CPlan(7, 512, 1)
    with pytest.raises(ValueError):
        CPlan(0, 500, 1)
    plan = CPlan(0, 128, 1, 4)
    img = np.zeros((1, 128, 128, 3), np.float32)
    with pytest.raises(ValueError):
        plan.detect_host(img)                            # weights not bound
    plan.close()
