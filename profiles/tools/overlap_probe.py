"""Timing probe: how much of the frozen-backbone forward of the NEXT batch hides behind the BiFPN / head forward +
backward of the current one (D0, batch 32, bf16)?  Captures the training plan as two graph segments (backbone |
rest) and replays them back to back on one stream vs. side by side on two."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from efficientdet_b200 import engine
from efficientdet_b200.model import efficientdet
from efficientdet_b200.optimizers import SGD
B, S = 32, 512
model = efficientdet(0, num_classes=20, dtype="bf16", image_size=S, just_training_model=True, seed=1)
model.freeze_backbone()
model.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9))
tr = model._trainer
plan = tr.plan(B, False)
nb = next(i for i, op in enumerate(plan.ops) if "BiFPN" in (op.name or ""))
print("ops", len(plan.ops), "backbone ops", nb)
plan.run(); torch.cuda.synchronize()
engine.Plan.capture(plan, bounds=[nb, len(plan.ops)])
g0, g1 = plan.segment_graphs
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def seq():
    g0.replay(); g1.replay()
def par():
    cur = torch.cuda.current_stream()
    sa.wait_stream(cur); sb.wait_stream(cur)
    with torch.cuda.stream(sa): g0.replay()
    with torch.cuda.stream(sb): g1.replay()
    cur.wait_stream(sa); cur.wait_stream(sb)
print("backbone alone %.3f ms" % timed(lambda: g0.replay()))
print("bifpn+heads fwd/bwd alone %.3f ms" % timed(lambda: g1.replay()))
print("sequential %.3f ms" % timed(seq))
print("side by side %.3f ms" % timed(par))
