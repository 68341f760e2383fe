import numpy as np, torch, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from util_model import perturb_weights, rel_err
from test_gpu_train import _targets
from efficientdet_b200.model import efficientdet
from efficientdet_b200.optimizers import SGD
from oracle import train as otrain
weighted = len(sys.argv)>1 and sys.argv[1]=="w"
size,C,B,phi=128,5,4,0
model=efficientdet(phi,num_classes=C,weighted_bifpn=weighted,image_size=size,dtype="fp32",drop_connect_rate=0,just_training_model=True)
W0=perturb_weights(model); model.freeze_backbone(); model.compile(optimizer=SGD(lr=0.01,decay=4e-5,momentum=0.9))
anchors,ann,reg_t,lab_t=_targets(size,B,C)
img=np.random.default_rng(5).standard_normal((B,size,size,3)).astype(np.float32)
print(model.train_on_batch(img,[reg_t,lab_t]))
fl,sl,grads,stats=otrain.loss_and_grads(W0,img,reg_t,lab_t,phi,C,weighted,False)
print(fl,sl)
for k,g in grads.items():
    got=model.net.grads[k].cpu().numpy(); sc=max(np.abs(g).max(),1e-12)
    e=np.abs(got-g).max()/sc
    print("%-50s %.2e %.2e %s"%(k,e,sc,"BAD" if e>2e-3 else ""))
