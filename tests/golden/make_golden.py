"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE's own code
(/root/reference/utils/anchors.py unmodified + its re-cythonized
utils/compute_overlap.pyx built by oracle/build_ref.py).

Only runs in the build container (needs /root/reference).  The only TF symbol
utils/anchors.py touches is keras.backend.floatx() (utils/anchors.py:18,49-51,93,95);
it is stubbed to return 'float32', TF's default.  utils/__init__.py imports cv2,
which is present.  Run:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402

build_ref.build_ref()
ref_overlap = build_ref.load_ref_compute_overlap()
assert ref_overlap is not None

tf = types.ModuleType("tensorflow")
k = types.ModuleType("tensorflow.keras")
b = types.ModuleType("tensorflow.keras.backend")
b.floatx = lambda: "float32"
k.backend = b
tf.keras = k
sys.modules.update({"tensorflow": tf, "tensorflow.keras": k, "tensorflow.keras.backend": b})
co = types.ModuleType("utils.compute_overlap")
co.compute_overlap = ref_overlap
sys.path.insert(0, "/root/reference")
import utils  # noqa: E402  (reference package)
sys.modules["utils.compute_overlap"] = co
utils.compute_overlap = co
from utils import anchors as ra  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


out = {}
# ---- generate_anchors
for s in (16, 32, 64, 128, 256, 512, 48):
    out["gen_%d" % s] = ra.generate_anchors(s, ra.AnchorParameters.default.ratios,
                                            ra.AnchorParameters.default.scales)
# ---- anchors_for_shape: full arrays for small shapes, digests for the model sizes
for shp in ((128, 128), (96, 160), (100, 150)):
    out["anchors_%dx%d" % shp] = ra.anchors_for_shape(shp)
meta = []
for S in (512, 640, 768, 896, 1024, 1280, 1408):
    a = ra.anchors_for_shape((S, S))
    assert a.dtype == np.float64
    meta.append((S, a.shape[0], float(a.sum()), sha(a), sha(a.astype(np.float32))))
    out["anchors_%d_first" % S] = a[0]
    out["anchors_%d_last" % S] = a[-1]
out["model_sizes"] = np.array([m[0] for m in meta])
out["model_counts"] = np.array([m[1] for m in meta])
out["model_sums"] = np.array([m[2] for m in meta])
out["model_sha_f64"] = np.array([m[3] for m in meta])
out["model_sha_f32"] = np.array([m[4] for m in meta])

# ---- compute_overlap (reference Cython)
rng = np.random.default_rng(11)
bx = rng.uniform(0, 400, (300, 2)); bw = rng.uniform(1, 200, (300, 2))
boxes = np.concatenate([bx, bx + bw], 1)
qx = rng.uniform(0, 400, (9, 2)); qw = rng.uniform(1, 200, (9, 2))
query = np.concatenate([qx, qx + qw], 1)
boxes[5] = query[2]                      # an exact match
boxes[6] = [500, 500, 510, 510]          # disjoint from everything
out["ov_boxes"], out["ov_query"] = boxes, query
out["ov_result"] = ref_overlap(boxes, query)
out["ov_kat"] = ref_overlap(np.array([[0, 0, 10, 10], [5, 5, 15, 15]], np.float64),
                            np.array([[0, 0, 10, 10]], np.float64))

# ---- bbox_transform
a128 = ra.anchors_for_shape((128, 128))
gt_rows = np.concatenate([rng.uniform(0, 60, (a128.shape[0], 2)),
                          rng.uniform(64, 128, (a128.shape[0], 2))], 1).astype(np.float32)
out["bt_gt"] = gt_rows
out["bt_result"] = ra.bbox_transform(a128, gt_rows)

# ---- anchor_targets_bbox: SURVEY Appendix B KAT (digest) + a small ragged batch (full)
a512 = ra.anchors_for_shape((512, 512))
reg, lab = ra.anchor_targets_bbox(
    a512, [np.zeros((512, 512, 3))],
    [{"bboxes": np.array([[100, 120, 300, 360], [10, 10, 60, 80]], np.float32),
      "labels": np.array([3, 7], np.float32)}], 20)
out["kat_pos_idx"] = np.nonzero(reg[0, :, 4] == 1)[0]
out["kat_ign_idx"] = np.nonzero(reg[0, :, 4] == -1)[0]
out["kat_reg_pos_rows"] = reg[0, reg[0, :, 4] == 1]
out["kat_reg_sum"] = np.array(reg[..., :4].astype(np.float64).sum())
out["kat_reg_sha"] = np.array(sha(reg))
out["kat_lab_sha"] = np.array(sha(lab))

ann = []
img_shapes = [(128, 128, 3), (100, 128, 3), (128, 90, 3), (128, 128, 3)]
for i, shp in enumerate(img_shapes):
    n = [3, 0, 5, 1][i]
    x1 = rng.uniform(0, 70, (n, 2)); wh = rng.uniform(12, 56, (n, 2))
    ann.append({"bboxes": np.concatenate([x1, x1 + wh], 1).astype(np.float32),
                "labels": rng.integers(0, 6, n).astype(np.float32)})
reg, lab = ra.anchor_targets_bbox(a128, [np.zeros(s) for s in img_shapes], ann, 6)
out["tg_img_shapes"] = np.array(img_shapes)
for i, an in enumerate(ann):
    out["tg_bboxes_%d" % i] = an["bboxes"]
    out["tg_labels_%d" % i] = an["labels"]
out["tg_regression"] = reg
out["tg_labels"] = lab

np.savez_compressed(os.path.join(HERE, "anchors_targets.npz"), **out)
print("wrote", os.path.join(HERE, "anchors_targets.npz"),
      os.path.getsize(os.path.join(HERE, "anchors_targets.npz")), "bytes")
for m in meta:
    print(m)
