"""Step rate of a COMPILED training plan replayed by the C library (effdet_replay_*; no Python in the loop besides the
ctypes call) for the headline workload: D0, batch 32, frozen backbone, bf16 -- to compare with bench.py's device-
resident rate of the Python-driven plan (same launches, same multi-lane capture order)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from efficientdet_b200.model import efficientdet
from efficientdet_b200.optimizers import SGD
from efficientdet_b200.plan import CReplay
from efficientdet_b200.plan_export import export_train_plan
B, S, C = 32, 512, 20
m = efficientdet(0, num_classes=C, dtype="bf16", image_size=S, just_training_model=True, seed=2024, drop_connect_rate=0)
m.freeze_backbone()
m.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9))
path = os.path.join(tempfile.mkdtemp(), "d0_train_b32.efd")
t0 = time.time()
info = export_train_plan(m, B, path, kmax=8)
print("exported %d launches, %d regions, %.1f MB in %.1f s" % (info["ops"], info["regions"], os.path.getsize(path) / 1e6, time.time() - t0))
del m
torch.cuda.empty_cache()
rp = CReplay(path)
rng = np.random.default_rng(0)
rp.write("images", rng.standard_normal((B, S, S, 3)).astype(np.float32))
boxes = np.zeros((B, 8, 4)); boxes[:, :3] = np.array([[40, 60, 200, 260], [300, 100, 480, 300], [100, 300, 260, 470]], np.float64)
rp.write("gt_boxes", boxes); rp.write("gt_labels", np.tile(np.array([1, 5, 7, 0, 0, 0, 0, 0], np.int32), (B, 1)))
rp.write("gt_counts", np.full((B,), 3, np.int32))
for i in range(5):
    rp.step(0.01)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    rp.step(0.01)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("C replay: %.3f ms/step = %.1f img/s; losses %s" % (ms, B / ms * 1e3, rp.read("losses", np.float32, (8,))[:2]))
