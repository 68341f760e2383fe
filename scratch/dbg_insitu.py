import numpy as np, torch, sys, os
os.environ["EFFDET_NO_REUSE"]="1"
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch.nn.functional as F
from util_model import perturb_weights
from test_gpu_train import _targets
from efficientdet_b200.model import efficientdet
from efficientdet_b200.optimizers import SGD
size,C,B,phi=128,5,4,0
model=efficientdet(phi,num_classes=C,image_size=size,dtype="fp32",drop_connect_rate=0,just_training_model=True)
W0=perturb_weights(model); model.freeze_backbone(); model.compile(optimizer=SGD(lr=0.0,momentum=0.0))
anchors,ann,reg_t,lab_t=_targets(size,B,C)
img=np.random.default_rng(5).standard_normal((B,size,size,3)).astype(np.float32)
model.train_on_batch(img,[reg_t,lab_t])
plan=list(model._trainer.plans.values())[0]
T=plan.tensor
for ht in plan.head_tape:
    layers=ht["layers"]
    for li in (2,1):
        L=layers[li]
        Wk=model.net.weights[L["name"]+"/kernel"]  # (3,3,ci,co)
        w=Wk.permute(3,2,0,1).contiguous()          # (co,ci,3,3)
        for l in range(5):
            dz=T(plan.gvals[id(L["ys"][l])]).permute(0,3,1,2)
            x=T(L["xs"][l])
            want=F.conv_transpose2d(dz,w,padding=1).permute(0,2,3,1)*(x>0)
            got=T(plan.gvals[id(L["xs"][l])])
            print(ht["scope"],li,l,tuple(x.shape),float((got-want).abs().max()/want.abs().max()))
print("---- head chain with torch autograd on GPU inputs")
torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False
net=model.net
for ht,dz_final,per in ((plan.head_tape[0],T(plan.dreg),4),(plan.head_tape[1],T(plan.dcls),C)):
    names=[L["name"] for L in ht["layers"]]+[ht["final"]]
    Ws={n:net.weights[n+"/kernel"].clone().double().requires_grad_(True) for n in names}
    Bs={n:net.weights[n+"/bias"].clone().double().requires_grad_(True) for n in names}
    outs=[]
    acts={}
    for l,f in enumerate(plan.pyramid):
        x=T(f).double().permute(0,3,1,2)
        for i,n in enumerate(names):
            x=F.conv2d(x,Ws[n].permute(3,2,0,1),Bs[n],padding=1)
            if i<len(names)-1:
                x=torch.relu(x); x.retain_grad(); acts[(n,l)]=x
        outs.append(x.permute(0,2,3,1).reshape(B,-1,per))
    out=torch.cat(outs,1)
    out.backward(dz_final.double())
    for n in names:
        g=net.grads[n+"/kernel"].double(); w=Ws[n].grad
        gb=net.grads[n+"/bias"].double(); wb=Bs[n].grad
        print(n, float((g-w).abs().max()/w.abs().max()), float((gb-wb).abs().max()/wb.abs().max()))
    # activation grads
    for i,L in enumerate(ht["layers"]):
        for l in range(5):
            got=T(plan.gvals[id(L["ys"][l])]).double().permute(0,3,1,2)
            want=acts[(L["name"],l)].grad
            print("  dz",L["name"],l,float((got-want).abs().max()/max(float(want.abs().max()),1e-30)))
