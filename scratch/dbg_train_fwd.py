import numpy as np, torch, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from util_model import perturb_weights, rel_err
from efficientdet_b200.model import efficientdet, EFFICIENTNET_DEPTHS
from efficientdet_b200 import train as T
from oracle import graph
size,C,B,phi=128,5,4,0
model=efficientdet(phi,num_classes=C,image_size=size,dtype="fp32",drop_connect_rate=0,just_training_model=True)
W0=perturb_weights(model)
for i in range(1,EFFICIENTNET_DEPTHS[phi]): model.layers[i].trainable=False
rng=np.random.default_rng(5)
img=rng.standard_normal((B,size,size,3)).astype(np.float32)
plan=T.TrainPlan(model.net,B,dense_labels=True)
plan.tensor(plan.images).copy_(torch.from_numpy(img).cuda())
plan.tensor(plan.reg_t).zero_(); plan.tensor(plan.lab_t).zero_()
stream=torch.cuda.current_stream().cuda_stream
for op in plan.ops[:plan.n_forward_ops]: op.fn(stream)
torch.cuda.synchronize()
taps={}
with torch.no_grad():
    r0,c0=graph.forward(W0,img,phi,C,False,taps=taps,bn_train_bifpn=True)
for name in ["C3","C4","C5"]+["BiFPN_%d_P%d"%(i,l) for i in range(2) for l in range(3,8)]:
    got=plan.tensor(plan.taps[name]).float().cpu().numpy()
    print(name, rel_err(got,taps[name].numpy()))
print("reg",rel_err(plan.tensor(plan.regression).cpu().numpy(),r0.numpy()),"cls",rel_err(plan.tensor(plan.classification).cpu().numpy(),c0.numpy()))
# per-record check of first layer laterals
for rec in plan.tape[:6]:
    print(rec["kind"],rec["name"],rec["rows"],rec.get("nblk"))
