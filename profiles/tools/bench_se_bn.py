"""Times effdet_se_bn_backward (fused SE + depthwise-BN/swish backward) and the separate pair
effdet_se_backward + effdet_bn_act_backward on D4 training shapes (bf16, batch 8).
usage: python profiles/tools/bench_se_bn.py            (EFFDET_B200_LIB selects another build)"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from efficientdet_b200 import _lib
lib = _lib.load()
dt = _lib.BF16
cases = {"b2b": (8, 256, 144, 6), "b3b": (8, 128, 336, 14), "b4b": (8, 64, 672, 28), "b5b": (8, 64, 960, 40),
         "b6b": (8, 32, 1632, 68), "b7b": (8, 32, 2688, 112)}
st = _lib.stream_ptr()
for name in (sys.argv[1:] or list(cases)):
    B, H, C, R = cases[name]
    HW = H * H
    z = torch.randn(B, HW, C, device="cuda").to(torch.bfloat16)
    dyg = (torch.randn(B, HW, C, device="cuda") * 0.1).to(torch.bfloat16)
    y = torch.randn(B, HW, C, device="cuda").to(torch.bfloat16)
    f = lambda *s: torch.randn(*s, device="cuda") * 0.2
    gamma, beta, mean, invstd = f(C) + 1, f(C), f(C), f(C).abs() + 0.5
    ua = gamma * invstd; ub = beta - mean * ua
    w1, b1, w2, b2 = f(C, R), f(R), f(R, C), f(C)
    sblk = lib.effdet_se_backward_blocks(HW, C, dt)
    se_sum = torch.randn(B, sblk, C, device="cuda"); gate = torch.rand(B, C, device="cuda")
    g = [torch.zeros_like(t) for t in (w1, b1, w2, b2, gamma, beta)]
    fcs = torch.empty(B * (2 * C * R + R + C), device="cuda"); dmean = torch.empty(B * C, device="cuda")
    dz = torch.empty_like(z); dy = torch.empty_like(z); k123 = torch.empty(3 * C, device="cuda")
    nb2 = lib.effdet_se_bn_backward_blocks(B, HW, C, dt)
    dgp2 = torch.empty(B * (nb2 + 1) * C, device="cuda"); bnp = torch.empty(B * (nb2 + 1) * 4 * C, device="cuda")
    bnr = torch.empty(B * 2 * C, device="cuda")
    dgb = lib.effdet_se_backward_blocks(HW, C, dt); dgp = torch.empty(B * dgb * C, device="cuda")
    rows = B * HW; nblk = lib.effdet_colreduce_blocks(rows, C, dt); part = torch.empty(2 * C * nblk, device="cuda")

    def fused():
        _lib.call("effdet_se_bn_backward", dyg.data_ptr(), z.data_ptr(), gate.data_ptr(), se_sum.data_ptr(), sblk,
                  w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), g[0].data_ptr(), g[1].data_ptr(),
                  g[2].data_ptr(), g[3].data_ptr(), gamma.data_ptr(), mean.data_ptr(), invstd.data_ptr(), ua.data_ptr(),
                  ub.data_ptr(), g[4].data_ptr(), g[5].data_ptr(), dz.data_ptr(), k123.data_ptr(), dgp2.data_ptr(),
                  bnp.data_ptr(), bnr.data_ptr(), nb2, fcs.data_ptr(), dmean.data_ptr(), B, HW, C, R, dt, st)

    def separate():
        _lib.call("effdet_se_backward", dyg.data_ptr(), y.data_ptr(), gate.data_ptr(), se_sum.data_ptr(), sblk,
                  w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), dy.data_ptr(), g[0].data_ptr(),
                  g[1].data_ptr(), g[2].data_ptr(), g[3].data_ptr(), dgp.data_ptr(), dgb, fcs.data_ptr(),
                  dmean.data_ptr(), B, HW, C, R, dt, st)
        _lib.call("effdet_bn_act_backward", dy.data_ptr(), z.data_ptr(), rows, C, gamma.data_ptr(), mean.data_ptr(),
                  invstd.data_ptr(), ua.data_ptr(), ub.data_ptr(), 0, _lib.ACT_SWISH, g[4].data_ptr(), g[5].data_ptr(),
                  dz.data_ptr(), k123.data_ptr(), part.data_ptr(), nblk, dt, st)
    out = []
    for fn in ([fused, separate] if hasattr(lib, "effdet_se_bn_backward") else [separate]):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            e0.record()
            for _ in range(5): fn()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 5)
        out.append(best * 1000)
    tb = B * HW * C * 2 / 1e6
    print("%-4s C=%4d HW=%6d tensor %6.1f MB  fused %7.1f us (%.2f TB/s over 5 passes)  separate %7.1f us" %
          (name, C, HW, tb, out[0], 5 * tb / out[0], out[-1]), flush=True)
