import numpy as np, torch, sys, ctypes
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from efficientdet_b200 import _lib
from oracle import graph
rng=np.random.default_rng(0)
B=3
for (cin,cout,Hs) in [(64,64,[8,4]),(64,36,[8,4]),(24,40,[6])]:
    w=(rng.standard_normal((3,3,cin,cout))/np.sqrt(9*cin)).astype(np.float32)
    wd=torch.from_numpy(w).cuda(); wt=torch.empty(9*cin*cout,device="cuda")
    _lib.call("effdet_conv_weight_transpose",wd.data_ptr(),wt.data_ptr(),9,cin,cout,_lib.stream_ptr())
    d=_lib.ConvDesc(); d.n_groups=len(Hs)
    keep=[]; wants=[]
    for i,H in enumerate(Hs):
        dz=rng.standard_normal((B,H,H,cout)).astype(np.float32)
        x=torch.tensor(rng.standard_normal((B,cin,H,H)),dtype=torch.float64,requires_grad=True)
        y=graph.conv2d(x,w.astype(np.float64),1)
        y.backward(torch.from_numpy(dz).permute(0,3,1,2).double())
        wants.append(x.grad.permute(0,2,3,1).numpy())
        dzd=torch.from_numpy(dz).cuda(); out=torch.full((B,H,H,cin),float("nan"),device="cuda")
        keep+= [dzd,out]
        d.x[i]=dzd.data_ptr(); d.y[i]=out.data_ptr(); d.H[i]=H; d.W[i]=H
    d.B,d.Cin,d.Cout,d.kh,d.kw,d.stride=B,cout,cin,3,3,1
    d.weight=wt.data_ptr(); d.act=0; d.in_dtype=d.out_dtype=0
    _lib.call("effdet_conv2d",ctypes.byref(d),_lib.stream_ptr()); torch.cuda.synchronize()
    for i in range(len(Hs)):
        got=keep[2*i+1].cpu().numpy()
        print(cin,cout,Hs[i], np.abs(got-wants[i]).max()/np.abs(wants[i]).max())
