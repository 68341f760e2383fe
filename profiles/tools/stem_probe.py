import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from efficientdet_b200.model import efficientdet
m, pm = efficientdet(0, num_classes=20, dtype="bf16", image_size=512, seed=1)
net = m.net
for u8 in (False, True):
    plan = net.plan(32, u8_input=u8)
    prof = plan.profile(iters=3)
    st = [o for o in prof if o['kind'] == 'stem']
    print('u8' if u8 else 'f32', [(o['name'], round(o['ms'], 4)) for o in st], 'total', round(sum(o['ms'] for o in prof), 3))
