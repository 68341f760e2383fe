"""Per-launch times of the D0 batch-1 inference plan (bf16), largest first, and per kind."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from efficientdet_b200.model import efficientdet
m, pm = efficientdet(0, num_classes=90, dtype="bf16", image_size=512, seed=1)
plan = m.net.plan(1)
prof = plan.profile(iters=5)
k = collections.Counter(); n = collections.Counter()
for o in prof: k[o['kind']] += o['ms']; n[o['kind']] += 1
print("total %.4f ms over %d launches" % (sum(o['ms'] for o in prof), len(prof)))
print([(a, n[a], round(b, 4)) for a, b in k.most_common()])
for o in prof:
    print("%-34s %-14s %.2f us" % (o['name'][:34], o['kind'], o['ms'] * 1e3))
