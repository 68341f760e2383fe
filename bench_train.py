"""Training legs of bench.py (kept OUTSIDE the efficientdet_b200 package: the cpu_baseline / reference-arm
functions execute oracle/, which only tests/, smoke() and bench.py may do).

  bench_train            our arm: device-resident step rate, end-to-end rate through Trainer.fit_prefetched
                         (float and raw uint8 host images), per-kind roofline
  cpu_baseline_train     bounded sample of the same training step on the host cores (oracle port)
  bench_train_reference  `bench.py --impl reference` for the training workloads
"""
import os
import time

import numpy as np
import torch

from efficientdet_b200 import _lib


def _synthetic_gt(B, S, C, seed):
    """n~U{1..8} boxes per image, w,h~U(32,256), fully inside, labels~U{0..C-1}."""
    rng = np.random.default_rng(seed)
    ann = []
    for _ in range(B):
        n = int(rng.integers(1, 9))
        wh = rng.uniform(32, min(256, S / 2), (n, 2))
        xy = rng.uniform(0, 1, (n, 2)) * (S - wh)
        ann.append({"bboxes": np.concatenate([xy, xy + wh], 1).astype(np.float32),
                    "labels": rng.integers(0, C, n).astype(np.float32)})
    return ann


def bench_train(args, rank, world, phi, B, C, dtype, weighted, dev, freeze_backbone=True, workload=None,
                sub_record=False):
    """sub_record: the secondary workload reported inside the headline line (no uint8 leg, no CPU baseline)."""
    workload = workload or args.workload
    import bench as bench_mod
    from efficientdet_b200.model import efficientdet
    from efficientdet_b200.optimizers import SGD
    from efficientdet_b200.utils.anchors import _pack_annotations, anchors_for_shape
    from efficientdet_b200.utils.tpu import tpu_focal, tpu_smooth_l1
    S = [512, 640, 768, 896, 1024, 1280, 1408][phi]
    hbm, tflops, peak_src = bench_mod.peaks()
    model = efficientdet(phi, num_classes=C, weighted_bifpn=weighted, dtype=dtype, drop_connect_rate=0,
                         just_training_model=True, seed=2024)
    if freeze_backbone:
        model.freeze_backbone()
    model.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9),
                  loss={"regression": tpu_smooth_l1(), "classification": tpu_focal(alpha=0.25, gamma=1.5)})
    tr = model._trainer
    plan = tr.plan(B, dense=False)
    anchors_d = torch.from_numpy(anchors_for_shape((S, S))).to(dev)
    n_sets = max(2, min(8, int(400e6 // (B * S * S * 12)) + 1))
    host_imgs = bench_mod.synth_images(B, S, n_sets, 1234 + rank)
    dev_imgs = [torch.from_numpy(h).to(dev) for h in host_imgs]
    pinned = [torch.from_numpy(h).pin_memory() for h in host_imgs]
    gts = []
    for i in range(n_sets):
        gt, gl, cnt, kmax = _pack_annotations(_synthetic_gt(B, S, C, 7 + 100 * rank + i))
        hw = np.tile(np.array([[float(S), float(S)]]), (B, 1))
        host = [torch.from_numpy(a).pin_memory() for a in (gt, gl, cnt, hw)]
        gts.append(dict(host=host, dev=[t.to(dev) for t in host], kmax=kmax))
    img_buf = plan.tensor(plan.images)

    def core(i, imgs_dev, gt_dev, kmax):
        img_buf.copy_(imgs_dev, non_blocking=True)
        tr.targets_into_plan(plan, anchors_d, gt_dev[0], gt_dev[1], gt_dev[2], gt_dev[3], kmax)
        tr.run_step(plan)

    def step_device(i):
        g = gts[i % n_sets]
        core(i, dev_imgs[i % n_sets], g["dev"], g["kmax"])

    step_device(0)                       # eager warm-up (sets kernel attributes)
    torch.cuda.synchronize(dev)
    plan.capture()
    warm = max(args.warmup, 3)
    for i in range(warm):
        step_device(i)
    torch.cuda.synchronize(dev)
    n0 = _lib.launch_count()
    step_device(0)
    torch.cuda.synchronize(dev)
    eager = _lib.launch_count() - n0
    launches_per_step = eager + sum(1 for op in plan.ops if op.kind not in ("memset",))

    def timed(fn, steps):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)
        st = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(steps):
            fn(i)
        e1.record(st)
        torch.cuda.synchronize(dev)
        from efficientdet_b200 import parallel
        return parallel.max_over_ranks(e0.elapsed_time(e1), dev)

    def run_e2e(steps):
        gen = ((pinned[i % n_sets], gts[i % n_sets]["host"], gts[i % n_sets]["kmax"]) for i in range(steps))
        for _ in tr.fit_prefetched(plan, anchors_d, gen):
            pass

    def timed_e2e(steps):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)
        st = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        run_e2e(steps)
        e1.record(st)
        torch.cuda.synchronize(dev)
        from efficientdet_b200 import parallel
        return parallel.max_over_ranks(e0.elapsed_time(e1), dev)

    # the plan is ~10^5 long-lived Python objects (ops, closures, descriptors): keep the cyclic collector from
    # walking them during the timed loops (a full collection is a 50-150 ms pause = several D4 steps)
    import gc
    gc.collect()
    gc.freeze()
    clocks = bench_mod.Clocks(dev.index) if rank == 0 else None
    ms = timed(step_device, args.steps)
    run_e2e(2)
    ms_e2e = timed_e2e(args.steps)
    clk = clocks.stop() if clocks else {}

    # the gradient all-reduce alone (the step's only collective), CUDA events, max over ranks
    from efficientdet_b200 import parallel as _par
    g_flat = model.net.grad_flat[0 if not freeze_backbone else model.net.backbone_end:]
    allreduce_us = None
    if world > 1:
        for _ in range(3):
            _par.allreduce_gradients_(g_flat)
        torch.distributed.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(dev))
        for _ in range(10):
            _par.allreduce_gradients_(g_flat)
        e1.record(torch.cuda.current_stream(dev))
        torch.cuda.synchronize(dev)
        allreduce_us = _par.max_over_ranks(e0.elapsed_time(e1), dev) / 10 * 1e3
        g_flat.zero_()
    ms_e2e8 = None
    if not sub_record:
        ms_e2e8 = _e2e_uint8(args, rank, world, tr, anchors_d, gts, n_sets, B, S, dev)
    losses = plan.tensor(plan.loss_out).cpu().numpy().tolist()

    prof = plan.profile(iters=3)
    if os.environ.get("EFFDET_DUMP_OPS"):
        import json
        with open(os.environ["EFFDET_DUMP_OPS"], "w") as f:
            json.dump(prof, f)
    by_kind = {}
    for r in prof:
        k2 = by_kind.setdefault(r["kind"], dict(ms=0.0, bytes=0, flops=0, n=0))
        k2["ms"] += r["ms"]; k2["bytes"] += r["bytes"]; k2["flops"] += r["flops"]; k2["n"] += 1
    total_ms = sum(v["ms"] for v in by_kind.values())
    dom_kind = max(by_kind, key=lambda k_: by_kind[k_]["ms"])
    dom = by_kind[dom_kind]
    ai = dom["flops"] / max(dom["bytes"], 1)
    if ai > tflops * 1e12 / (hbm * 1e9):
        roof = dict(bound="tensor", achieved=dom["flops"] / (dom["ms"] * 1e-3) / 1e12, peak=tflops,
                    unit="TFLOP/s")
    else:
        roof = dict(bound="hbm", achieved=dom["bytes"] / (dom["ms"] * 1e-3) / 1e9, peak=hbm, unit="GB/s")
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof.update(traffic=bench_mod.measured_traffic(workload, dom_kind), kernel=dom_kind, launches=dom["n"], share_of_step=dom["ms"] / total_ms,
                peak_source=peak_src,
                per_kind_ms={k_: round(v["ms"], 4) for k_, v in sorted(by_kind.items())})
    # the CPU leg runs on rank 0 at N=1 only (at N>1 the other ranks would spin in a barrier for its 12 s)
    cpu = cpu_baseline_train(phi, C, weighted, S, freeze_backbone=freeze_backbone) \
        if (rank == 0 and world == 1 and not sub_record and not os.environ.get('EFFDET_BENCH_NO_CPU')) else None
    imgs = B * world * args.steps
    h2d = B * S * S * 3 * 4 + sum(t.numel() * t.element_size() for t in gts[0]["host"])
    grad_mb = 4 * g_flat.numel() / 1e6
    out = {
        "metric": "images/sec", "value": imgs / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype,
        "data": "synthetic (numpy default_rng images + VOC-shaped boxes, random-init weights)",
        "config": {"workload": workload, "phi": phi, "image_size": S, "batch_per_gpu": B,
                   "global_batch": B * world, "num_classes": C, "weighted_bifpn": weighted,
                   "freeze_backbone": freeze_backbone, "optimizer": "SGD(lr=.01, decay=4e-5, momentum=.9)",
                   "step": "device anchor targets -> forward (BN batch stats in %s) -> focal + "
                           "smooth-L1 -> backward (%s) -> grad all-reduce -> SGD"
                           % (("BiFPN", "heads + BiFPN") if freeze_backbone else
                              ("every layer", "heads + BiFPN + backbone")),
                   "l2": "inputs rotate over %d image sets (%.0f MB) > 126 MB L2; activations %.0f MB"
                         % (n_sets, n_sets * B * S * S * 12 / 1e6, plan.activation_bytes / 1e6),
                   "parallelism": "dp%d (NCCL all-reduce of %.1f MB fp32 gradients)" % (world, grad_mb),
                   "final_losses": losses[:2]},
        "allreduce": {"mb": grad_mb, "us": allreduce_us,
                      "note": "the flat trainable-gradient range, timed alone (10 calls, CUDA events, max over "
                              "ranks); null at 1 GPU"},
        "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": 32},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roof, "cpu_baseline": cpu, "clocks": clk,
    }
    if ms_e2e8 is not None:
        out["e2e_uint8"] = {"value": imgs / (ms_e2e8 * 1e-3), "unit": "images/s",
                            "h2d_bytes_per_step": int(h2d - B * S * S * 9), "d2h_bytes_per_step": 32,
                            "note": "same loop, raw letterboxed uint8 RGB input (train_tpu.py TFRecord format; "
                                    "normalize_image on the device)"}
    if sub_record:
        for k in ("metric", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling", "vs_baseline",
                  "data", "cpu_baseline", "clocks"):
            out.pop(k, None)
    # free this workload's plans before the next one is built
    del plan, tr
    model._trainer = None
    torch.cuda.empty_cache()
    return out


def _e2e_uint8(args, rank, world, tr, anchors_d, gts, n_sets, B, S, dev):
    """The same end-to-end loop fed with raw letterboxed uint8 images (what train_tpu.py:170-183 decodes from the
    TFRecord PNGs): normalize_image runs on the device, the image upload is 3 B/pixel.  -> ms for args.steps."""
    import gc
    rng8 = np.random.default_rng(4321 + rank)
    pinned8 = [torch.from_numpy(rng8.integers(0, 256, (B, S, S, 3), dtype=np.uint8)).pin_memory()
               for _ in range(n_sets)]
    plan8 = tr.plan(B, dense=False, u8=True)
    plan8.tensor(plan8.input_images).copy_(pinned8[0].to(dev))
    g0 = gts[0]
    tr.targets_into_plan(plan8, anchors_d, g0["dev"][0], g0["dev"][1], g0["dev"][2], g0["dev"][3], g0["kmax"])
    plan8.run()
    tr.apply_gradients()
    torch.cuda.synchronize(dev)
    plan8.capture()

    def run_e2e8(steps):
        gen = ((pinned8[i % n_sets], gts[i % n_sets]["host"], gts[i % n_sets]["kmax"]) for i in range(steps))
        for _ in tr.fit_prefetched(plan8, anchors_d, gen):
            pass
    run_e2e8(3)
    gc.collect()
    gc.freeze()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream(dev))
    run_e2e8(args.steps)
    e1.record(torch.cuda.current_stream(dev))
    torch.cuda.synchronize(dev)
    from efficientdet_b200 import parallel as _par
    return _par.max_over_ranks(e0.elapsed_time(e1), dev)


def _cpu_train_step_fn(phi, C, weighted, S, batch, freeze_backbone=True):
    import bench as bench_mod
    from oracle import anchors as oa, train as otrain
    W = bench_mod._random_weights(phi, C, weighted)
    anchors = oa.anchors_for_shape((S, S))
    img = bench_mod.synth_images(batch, S, 1, 99)[0]
    ann = _synthetic_gt(batch, S, C, 7)
    vel = {}

    def step():
        from oracle import overlap_c
        reg_t, lab_t = overlap_c.anchor_targets_bbox(anchors, [(S, S, 3)] * batch, ann, C)
        _, _, grads, _ = otrain.loss_and_grads(W, img, reg_t, lab_t, phi, C, weighted, False,
                                               dtype=torch.float32, freeze_backbone=freeze_backbone)
        otrain.sgd_step(W, grads, vel)
    return step


def cpu_baseline_train(phi, C, weighted, S, budget_s=12.0, batch=2, freeze_backbone=True):
    """Bounded sample of the same training step on the host cores: one untimed warm-up step, then whole
    steps until ~budget_s seconds of CPU work have been measured."""
    torch.set_num_threads(os.cpu_count())
    step = _cpu_train_step_fn(phi, C, weighted, S, batch, freeze_backbone)
    step()                                   # warm-up (allocator, oneDNN primitive caches)
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += batch
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "%d image(s): the same training step (C target assignment + torch-CPU fp32 autograd "
                      "of the reference graph, %s, SGD) in batches of %d, %.1f s"
                      % (n, "frozen backbone" if freeze_backbone else "nothing frozen", batch, dt)}


def bench_train_reference(args, phi, B, C, weighted, freeze_backbone=True):
    S = [512, 640, 768, 896, 1024, 1280, 1408][phi]
    torch.set_num_threads(os.cpu_count())
    # the workload's own batch and the requested step counts when the run fits ~4 minutes (D0, batch 32: ~2.5 s
    # per step on 16 cores); else whole steps of a smaller batch
    budget_s = 240.0
    probe = _cpu_train_step_fn(phi, C, weighted, S, 2, freeze_backbone)
    probe()                                  # allocator / oneDNN primitive caches
    t0 = time.perf_counter()
    probe()
    per_img = (time.perf_counter() - t0) / 2
    warm = max(0, min(args.warmup, 1))
    batch = B
    while batch > 1 and per_img * batch * (args.steps + warm) > budget_s:
        batch //= 2
    steps = args.steps
    while steps > 1 and per_img * batch * (steps + warm) > budget_s:
        steps -= 1
    step = _cpu_train_step_fn(phi, C, weighted, S, batch, freeze_backbone)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    v = batch * steps / dt
    return {
        "impl": "reference", "metric": "images/sec", "value": v, "unit": "images/s", "n_gpus": 1,
        "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (numpy default_rng images + VOC-shaped boxes, random-init weights)",
        "config": {"workload": args.workload, "phi": phi, "image_size": S, "batch_per_gpu": B,
                   "num_classes": C, "weighted_bifpn": weighted, "freeze_backbone": freeze_backbone,
                   "reference_batch": batch},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": "%d-image training steps on torch-CPU fp32 (reference graph restated; "
                                   "TensorFlow not installable)" % batch},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
