"""Data-parallel training step on GPUs: the bucketed, overlapped gradient all-reduce.
* 1 GPU: the segmented-graph + second-stream path (forced) is bit-identical to the single-graph path.
* 2 GPUs (NCCL; skipped when fewer are visible -- run with `gpurun --gpus 2`): after one step both replicas hold
  identical weights, equal to W0 - lr * mean over replicas of the per-shard gradients, each shard's loss
  normalised by ITS OWN number of positive anchors and BatchNorm statistics taken per replica
  (tf.distribute.MirroredStrategy semantics, utils/tpu.py:64-66,148-151) -- checked against the
  fp64 autograd oracle run per shard."""
import os
import socket

import numpy as np
import pytest
import torch

from util_model import golden_weight, rel_err, rel_l2

pytestmark = pytest.mark.gpu


def _targets(size, B, C, seed):
    from oracle import anchors as oa
    rng = np.random.default_rng(seed)
    anchors = oa.anchors_for_shape((size, size))
    ann = []
    for _ in range(B):
        n = int(rng.integers(1, 6))
        wh = rng.uniform(size * 0.1, size * 0.5, (n, 2))
        xy = rng.uniform(0, 1, (n, 2)) * (size - wh)
        ann.append({"bboxes": np.concatenate([xy, xy + wh], 1).astype(np.float32),
                    "labels": rng.integers(0, C, n).astype(np.float32)})
    return oa.anchor_targets_bbox(anchors, [(size, size, 3)] * B, ann, C)


def _model(phi, C, size, dtype, freeze_backbone, seed=5):
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.optimizers import SGD
    model = efficientdet(phi, num_classes=C, weighted_bifpn=True, image_size=size, dtype=dtype,
                         drop_connect_rate=0, just_training_model=True)
    W = {k: golden_weight(k, v.shape, seed) for k, v in model.get_weights_dict().items()}
    model.set_weights_dict(W, strict=True)
    if freeze_backbone:
        model.freeze_backbone()
    model.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9))
    return model, W


@pytest.mark.parametrize("freeze_backbone", [True, False])
def test_bucketed_step_is_bit_identical_on_one_gpu(freeze_backbone, monkeypatch):
    size, C, B = 256, 5, 4
    reg_t, lab_t = _targets(size, B, C, 7)
    img = np.random.default_rng(5).standard_normal((B, size, size, 3)).astype(np.float32)
    res = []
    for forced in (False, True):
        if forced:
            monkeypatch.setenv("EFFDET_FORCE_BUCKETS", "1")
            monkeypatch.setenv("EFFDET_BUCKET_MIN_ELEMS", "1000")
        model, _ = _model(0, C, size, "bf16", freeze_backbone)
        for _ in range(2):
            loss = model.train_on_batch(img, [reg_t, lab_t])
        plan = list(model._trainer.plans.values())[0]
        n_seg = len(plan.segment_graphs)
        assert n_seg == ((2 if freeze_backbone else 4) if forced else 1), n_seg
        res.append((loss, model.get_weights_dict()))
    assert res[0][0] == res[1][0]
    for k in res[0][1]:
        assert np.array_equal(res[0][1][k], res[1][1][k]), k


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), EFFDET_BUCKET_MIN_ELEMS="1000")
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    size, C, B = 256, 5, 4
    model, W0 = _model(0, C, size, "fp32", False)
    rng = np.random.default_rng(50 + rank)                  # this replica's shard of the global batch
    img = rng.standard_normal((B, size, size, 3)).astype(np.float32)
    reg_t, lab_t = _targets(size, B, C, 70 + rank)
    loss = model.train_on_batch(img, [reg_t, lab_t])
    plan = list(model._trainer.plans.values())[0]
    W1 = model.get_weights_dict()
    G = {k: v.cpu().numpy() / world for k, v in model.net.grads.items()}     # all-reduced in place: SUM
    q.put((rank, loss, len(plan.segment_graphs), W1, G))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_replica_step_matches_oracle_mean_of_shard_gradients():
    import torch.multiprocessing as mp
    from oracle import train as otrain
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, loss0, nseg0, Wa, Ga), (_, loss1, nseg1, Wb, Gb) = got
    assert nseg0 == nseg1 == 4                              # heads | BiFPN | stages 7-5 | stages 4-1 + stem
    for k in Wa:                                            # replicas stay in lock step (BN moving stats are
        if not k.endswith(("moving_mean", "moving_variance")):     # per replica by design)
            assert np.array_equal(Wa[k], Wb[k]), k
    size, C, B, phi = 256, 5, 4, 0
    from efficientdet_b200.model import efficientdet      # only for the weight manifest
    W0 = {k: golden_weight(k, v.shape, 5) for k, v in Wa.items()}
    grads = []
    for rank, loss in ((0, loss0), (1, loss1)):
        img = np.random.default_rng(50 + rank).standard_normal((B, size, size, 3)).astype(np.float32)
        reg_t, lab_t = _targets(size, B, C, 70 + rank)
        fl, sl, g, _ = otrain.loss_and_grads(W0, img, reg_t, lab_t, phi, C, True, False, freeze_backbone=False)
        assert abs(loss[2] - fl) / fl < 2e-4 and abs(loss[1] - sl) / max(sl, 1e-9) < 2e-4   # local normalisers
        grads.append(g)
    bad = {}
    for k in grads[0]:
        gm = 0.5 * (grads[0][k] + grads[1][k])
        if np.abs(gm).max() < 1e-12 or k.startswith("w_bi_fpn_add"):
            continue
        assert np.array_equal(Ga[k], Gb[k]), k              # both replicas hold the same reduced gradient
        e = rel_l2(Ga[k], gm)
        if not e < 8e-2:
            bad[k] = e
    assert not bad, bad
    for k in ("stem_conv/kernel", "block5a_expand_conv/kernel", "BiFPN_1_P4_conv/kernel",
              "class_head/pyramid_classification/kernel"):      # one key per bucket: SGD saw the MEAN gradient
        want = W0[k].astype(np.float64) - 0.01 * Ga[k]
        assert rel_err(Wa[k], want) < 1e-5, k
