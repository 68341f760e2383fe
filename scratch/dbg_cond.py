import numpy as np, torch, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import train as otrain, graph, anchors as oa
import bench
size,C,B,phi=256,5,8,0
W=bench._random_weights(phi,C,False)
rng=np.random.default_rng(2024)
for k,v in W.items():
    if k.endswith("/gamma"): W[k]=rng.uniform(.5,1.5,v.shape).astype(np.float32)
    elif k.endswith("/beta") or k.endswith("/moving_mean"): W[k]=rng.normal(0,.1,v.shape).astype(np.float32)
    elif k.endswith("/moving_variance"): W[k]=rng.uniform(.5,1.5,v.shape).astype(np.float32)
    elif k.startswith(("box_head","class_head")) and k.endswith("/kernel"): W[k]=(rng.standard_normal(v.shape)*np.sqrt(2/(9*v.shape[2]))).astype(np.float32)
from test_gpu_train import _targets
anchors,ann,reg_t,lab_t=_targets(size,B,C)
img=np.random.default_rng(5).standard_normal((B,size,size,3)).astype(np.float32)
r64=otrain.loss_and_grads(W,img,reg_t,lab_t,phi,C,False,False,dtype=torch.float64)
r32=otrain.loss_and_grads(W,img,reg_t,lab_t,phi,C,False,False,dtype=torch.float32)
print(r64[0],r32[0],r64[1],r32[1])
for k in list(r64[2])[:6]+list(r64[2])[60:66]+[k for k in r64[2] if "head" in k]:
    a,b=r64[2][k],r32[2][k]
    print("%-50s max %.2e  l2 %.2e"%(k,np.abs(a-b).max()/np.abs(a).max(), np.linalg.norm(a-b)/np.linalg.norm(a)))
