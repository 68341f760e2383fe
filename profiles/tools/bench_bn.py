import sys, ctypes, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from efficientdet_b200 import _lib
lib=_lib.load()
cases={"b2":(8*256*256,144),"b3":(8*128*128,192),"b5":(8*64*64,672),"b6":(8*32*32,960),"b7":(8*32*32,1632)}
for name in (sys.argv[1:] or list(cases)):
    rows,C=cases[name]
    z=torch.randn((rows,C),device="cuda").to(torch.bfloat16); dy=torch.randn((rows,C),device="cuda").to(torch.bfloat16)
    dz=torch.empty_like(z); y=torch.empty_like(z)
    gamma=torch.rand(C,device="cuda")+0.5; beta=torch.randn(C,device="cuda")*0.1
    mm=torch.zeros(C,device="cuda"); mv=torch.ones(C,device="cuda")
    sc=torch.empty(C,device="cuda"); sh=torch.empty(C,device="cuda"); sm=torch.empty(C,device="cuda"); si=torch.empty(C,device="cuda")
    nblk=lib.effdet_colreduce_blocks(rows,C,_lib.BF16)
    part=torch.empty(2*C*nblk,device="cuda"); k123=torch.empty(3*C,device="cuda"); dg=torch.empty(C,device="cuda"); db=torch.empty(C,device="cuda")
    st=_lib.stream_ptr()
    def stats(): _lib.call("effdet_bn_train_stats",z.data_ptr(),rows,C,gamma.data_ptr(),beta.data_ptr(),1e-3,0.99,mm.data_ptr(),mv.data_ptr(),sc.data_ptr(),sh.data_ptr(),sm.data_ptr(),si.data_ptr(),part.data_ptr(),nblk,_lib.BF16,st)
    def apply(): _lib.call("effdet_scale_shift_act",z.data_ptr(),sc.data_ptr(),sh.data_ptr(),y.data_ptr(),rows,C,_lib.ACT_SWISH,_lib.BF16,st)
    def bwd(): _lib.call("effdet_bn_act_backward",dy.data_ptr(),z.data_ptr(),rows,C,gamma.data_ptr(),sm.data_ptr(),si.data_ptr(),sc.data_ptr(),sh.data_ptr(),0,_lib.ACT_SWISH,dg.data_ptr(),db.data_ptr(),dz.data_ptr(),k123.data_ptr(),part.data_ptr(),nblk,_lib.BF16,st)
    for f,nm,byt in ((stats,"stats",rows*C*2),(apply,"apply",rows*C*4),(bwd,"bwd",rows*C*10)):
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/10
        print("%s rows=%d C=%d nblk=%d %-6s %.4f ms %7.1f GB/s"%(name,rows,C,nblk,nm,ms,byt/ms/1e6))
