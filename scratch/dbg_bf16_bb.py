import sys, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from test_gpu_train import _targets
from util_model import perturb_weights, rel_l2
from efficientdet_b200.model import efficientdet
from efficientdet_b200.optimizers import SGD
size, C, B, phi = 256, 5, 4, 0
anchors, ann, reg_t, lab_t = _targets(size, B, C)
img = np.random.default_rng(5).standard_normal((B, size, size, 3)).astype(np.float32)
res = {}
for dt in ("fp32", "bf16"):
    model = efficientdet(phi, num_classes=C, weighted_bifpn=False, image_size=size, dtype=dt,
                         drop_connect_rate=0, just_training_model=True)
    perturb_weights(model)
    model.compile(optimizer=SGD(lr=0.01, decay=4e-5, momentum=0.9))
    loss = model.train_on_batch(img, [reg_t, lab_t])
    res[dt] = (loss, {k: v.cpu().numpy().copy() for k, v in model.net.grads.items()})
print(res["fp32"][0], res["bf16"][0])
for k in res["fp32"][1]:
    a, b = res["bf16"][1][k], res["fp32"][1][k]
    if np.abs(b).max() < 1e-12: continue
    if k.endswith(("moving_mean","moving_variance")): continue
    print("%-48s %.3f  |bf16| %.3e |fp32| %.3e" % (k, rel_l2(a, b), np.linalg.norm(a), np.linalg.norm(b)))
