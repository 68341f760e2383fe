import numpy as np, torch, sys
sys.path.insert(0,'.')
from oracle import losses
from efficientdet_b200 import _lib
B,N,C=2,3000,7
rng=np.random.default_rng(0)
p=rng.uniform(0.001,0.999,(B,N,C)).astype(np.float32)
reg=rng.normal(0,1.2,(B,N,4)).astype(np.float32)
state=rng.choice([-1,0,1],(B,N),p=[0.1,0.8,0.1]).astype(np.float32)
clsid=rng.integers(0,C,(B,N))
lab=np.zeros((B,N,C+1),np.float32); bi,ni=np.nonzero(state==1); lab[bi,ni,clsid[bi,ni]]=1; lab[...,C]=state
reg_t=np.concatenate([rng.normal(0,1,(B,N,4)),state[...,None]],-1).astype(np.float32)
pt=torch.tensor(p,dtype=torch.float64,requires_grad=True)
fl=losses.focal(torch.tensor(lab,dtype=torch.float64),pt,0.25,1.5); fl.backward()
want=(pt.grad*pt.detach()*(1-pt.detach())).numpy()
lib=_lib.load()
d=lambda a,dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to("cuda",dt)
pd,rd,rtd,labd=d(p),d(reg),d(reg_t),d(lab)
dcls=torch.empty((B,N,C),device="cuda"); dreg=torch.empty((B,N,4),device="cuda"); out8=torch.zeros(8,device="cuda")
wsb=lib.effdet_detection_losses_workspace_size(); ws=torch.empty(wsb,dtype=torch.uint8,device="cuda")
st=d(state,torch.int8); cl=d(np.where(state==1,clsid,-1),torch.int32)
for dense in (True,False):
    _lib.call("effdet_detection_losses",pd.data_ptr(),rd.data_ptr(),rtd.data_ptr(),labd.data_ptr() if dense else None,st.data_ptr(),cl.data_ptr(),B,N,C,0.25,1.5,1.0,1.0,dcls.data_ptr(),dreg.data_ptr(),out8.data_ptr(),ws.data_ptr(),wsb,None,None,None,0,0,0,_lib.stream_ptr())
    g=dcls.cpu().numpy()
    err=np.abs(g-want); i=np.unravel_index(err.argmax(),g.shape)
    print(dense, out8.cpu().numpy(), err.max(), np.abs(want).max(), i, g[i], want[i], p[i], lab[i[0],i[1]], )
    bad=np.argwhere(err>1e-6*np.abs(want).max()); print(len(bad), bad[:10].tolist())
