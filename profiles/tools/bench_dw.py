import sys, ctypes, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from efficientdet_b200 import _lib
cases={"b2a":(32,256,96,3,2),"b2b":(32,128,144,3,1),"b3a":(32,128,144,5,2),"b3b":(32,64,240,5,1),"b4b":(32,32,480,3,1),
       "b5b":(32,32,672,5,1),"b6a":(32,32,672,5,2),"b6b":(32,16,1152,5,1),"b1a":(32,256,32,3,1),"b7a":(32,16,1152,3,1)}
names=sys.argv[1:] or list(cases)
lib=_lib.load()
for name in names:
    B,H,C,k,s=cases[name]
    Ho=(H+s-1)//s
    x=torch.randn((B,H,H,C),device="cuda").to(torch.bfloat16)
    w=torch.randn((k,k,C),device="cuda")*0.2
    sc=torch.ones(C,device="cuda"); sh=torch.zeros(C,device="cuda")
    y=torch.empty((B,Ho,Ho,C),device="cuda",dtype=torch.bfloat16)
    nblk=lib.effdet_dwconv_se_blocks(B,H,H,C,s,_lib.BF16)
    part=torch.empty((B,nblk,C),device="cuda")
    st=_lib.stream_ptr()
    f=lambda: _lib.call("effdet_dwconv",x.data_ptr(),w.data_ptr(),sc.data_ptr(),sh.data_ptr(),y.data_ptr(),part.data_ptr(),nblk,B,H,H,C,k,s,_lib.ACT_SWISH,_lib.BF16,st)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    byt=B*(H*H+Ho*Ho)*C*2
    # reference check
    xr=x.float().permute(0,3,1,2)
    pt=max((Ho-1)*s+k-H,0); p0=pt//2; p1=pt-p0
    xr=torch.nn.functional.pad(xr,(p0,p1,p0,p1))
    yr=torch.nn.functional.conv2d(xr,w.permute(2,0,1)[:,None],stride=s,groups=C)
    yr=(yr*torch.sigmoid(yr)).permute(0,2,3,1)
    err=(y.float()-yr).abs().max().item()/yr.abs().max().item()
    se=(part.sum(1)-yr.sum((1,2))).abs().max().item()/yr.sum((1,2)).abs().max().item()
    print("%-4s B%d H%d C%d k%d s%d: %.4f ms %7.1f GB/s  err %.2e se_err %.2e"%(name,B,H,C,k,s,ms,byt/ms/1e6,err,se))
