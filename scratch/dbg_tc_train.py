import numpy as np, torch, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from util_model import perturb_weights, rel_l2
from test_gpu_train import _targets
from efficientdet_b200.model import efficientdet
from efficientdet_b200.optimizers import SGD
size,C,B,phi=256,5,4,0
anchors,ann,reg_t,lab_t=_targets(size,B,C)
img=np.random.default_rng(5).standard_normal((B,size,size,3)).astype(np.float32)
res={}
for tag,kw in (("simt",dict(tensor_cores=False)),("tc",dict(tensor_cores=True)),("simt2",dict(tensor_cores=False,seed=2024))):
    model=efficientdet(phi,num_classes=C,weighted_bifpn=True,image_size=size,dtype="bf16",drop_connect_rate=0,just_training_model=True,**kw)
    perturb_weights(model); model.freeze_backbone(); model.compile(optimizer=SGD(lr=0.01,decay=4e-5,momentum=0.9))
    if tag=="simt2":
        # perturb the input by 1 bf16 ulp-ish noise to gauge the sensitivity of bf16 gradients
        img2=img*(1+1e-3*np.random.default_rng(1).standard_normal(img.shape).astype(np.float32))
        loss=model.train_on_batch(img2,[reg_t,lab_t])
    else:
        loss=model.train_on_batch(img,[reg_t,lab_t])
    res[tag]=(loss,{k:v.cpu().numpy().copy() for k,v in model.net.grads.items() if k.startswith(("BiFPN_","box_head","class_head","w_bi"))})
print({k:v[0] for k,v in res.items()})
for k,g in res["simt"][1].items():
    if k.endswith(("moving_mean","moving_variance")) or np.abs(g).max()<1e-10: continue
    print("%-50s tc %.3f   noise %.3f"%(k,rel_l2(res["tc"][1][k],g),rel_l2(res["simt2"][1][k],g)))
