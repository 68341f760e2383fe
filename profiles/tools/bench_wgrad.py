import sys, ctypes, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from efficientdet_b200 import _lib
lib=_lib.load()
cases={"d0trunk":(32,[64,32,16,8,4],64,64,64),"d0cls":(32,[64,32,16,8,4],64,180,184),"d0box":(32,[64,32,16,8,4],64,36,40),
       "d4trunk":(8,[128,64,32,16,8],224,224,224),"d4cls":(8,[128,64,32,16,8],224,810,816),"d4box":(8,[128,64,32,16,8],224,36,40),
       "d2trunk":(16,[96,48,24,12,6],112,112,112)}
for name in (sys.argv[1:] or list(cases)):
    B,sizes,cin,cout,ldz=cases[name]
    d=_lib.WgradDesc(); d.n_groups=len(sizes); keep=[]
    byt=0
    for i,H in enumerate(sizes):
        x=torch.randn((B,H,H,cin),device="cuda").to(torch.bfloat16); dz=torch.randn((B,H,H,ldz),device="cuda").to(torch.bfloat16)
        keep+=[x,dz]; byt+=x.numel()*2+dz.numel()*2
        d.x[i],d.dz[i],d.H[i],d.W[i],d.dz_ld[i]=x.data_ptr(),dz.data_ptr(),H,H,ldz
    d.B,d.Cin,d.Cout,d.kh,d.kw,d.stride=B,cin,cout,3,3,1
    d.x_dtype=d.dz_dtype=_lib.BF16
    ns=lib.effdet_conv_wgrad_tc_splits(ctypes.byref(d))
    part=torch.empty(ns*(9*cin*cout+cout),device="cuda"); out=torch.empty((3,3,cin,cout),device="cuda"); db=torch.empty(cout,device="cuda")
    d.dweight,d.partial,d.n_splits,d.accumulate=out.data_ptr(),part.data_ptr(),ns,0
    d.dbias=db.data_ptr() if lib.effdet_conv_wgrad_tc_fuses_bias(ctypes.byref(d)) else None
    st=_lib.stream_ptr()
    f=lambda: _lib.call("effdet_conv_wgrad_tc",ctypes.byref(d),st)
    for _ in range(3): f()
    torch.cuda.synchronize()
    g=torch.cuda.CUDAGraph()
    s2=torch.cuda.Stream()
    with torch.cuda.graph(g,stream=s2):
        for _ in range(8): _lib.call("effdet_conv_wgrad_tc",ctypes.byref(d),torch.cuda.current_stream().cuda_stream)
    g.replay(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/24
    fl=sum(2*B*H*H*cin*cout*9 for H in sizes)
    print("%s splits=%d %.4f ms (wgrad + reduce) %.1f GB/s %.1f TF/s"%(name,ns,ms,byt/ms/1e6,fl/ms/1e9))
