"""Static launch plan for the EfficientDet forward pass on one GPU.

model.efficientdet() describes the network (Keras layer names, shapes); this module lowers it
for a fixed (batch, dtype) into a flat list of C-ABI kernel launches over preallocated NHWC
buffers (liveness-based reuse), optionally captured in a CUDA graph.  torch provides device
memory, streams and graph capture only -- every launch is a kernel of libeffdet_b200.so.

Replaces the execution of the Keras graph built by the reference's model.py:356-452 /
efficientnet.py:309-470 (call stack SURVEY.md section 3 (B)).
"""
import ctypes

import numpy as np
import os

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SWISH, BF16, F32

TORCH_DTYPE = {F32: torch.float32, BF16: torch.bfloat16, "i8": torch.int8, "i32": torch.int32,
               "u8": torch.uint8}
BN_EPS_BACKBONE = 1e-3            # keras default (efficientnet.py:233-236 passes no epsilon)
BN_EPS_BIFPN = 1e-4               # model.py:42-45
BN_MOMENTUM_BIFPN = 0.997
BN_MOMENTUM_BACKBONE = 0.99


class Val:
    """A planned activation tensor (NHWC)."""
    __slots__ = ("shape", "dtype", "name", "keep", "t", "last_use")

    def __init__(self, shape, dtype, name=None, keep=False):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = dtype
        self.name = name
        self.keep = keep          # never recycled (outputs, taps)
        self.t = None
        self.last_use = -1

    @property
    def nbytes(self):
        n = 1
        for s in self.shape:
            n *= s
        return n * {BF16: 2, F32: 4, "i8": 1, "i32": 4, "u8": 1}[self.dtype]

    @property
    def ptr(self):
        return self.t.data_ptr()


class Op:
    """One kernel launch.  `make` is called after buffers are assigned and returns a
    zero-argument callable taking only the stream pointer."""

    def __init__(self, kind, inputs, outputs, make, name="", nbytes=0, flops=0):
        self.kind, self.inputs, self.outputs, self.make, self.name = kind, inputs, outputs, make, name
        self.fn = None
        self.nbytes, self.flops = nbytes, flops     # algorithmic HBM bytes / FLOPs of the launch


def _call(name, *args):
    fn = getattr(_lib.load(), name)

    def run(stream):
        rc = fn(*args, stream)
        if rc != 0:
            _lib.check(rc)
    run.call = (name, args)          # what plan_export.py serialises: entry point + arguments (stream excluded)
    return run


class Plan:
    def __init__(self, net, batch, reuse_buffers=True, keep_taps=False, u8_input=False):
        self.net = net
        self.u8_input = bool(u8_input)     # images are raw letterboxed uint8 RGB; normalised inside the stem
        self.B = int(batch)
        self.dev = net.device
        self.dtype = net.dtype
        self.ops = []
        self.vals = []
        self.taps = {}
        self.keep_taps = keep_taps
        self.reuse = reuse_buffers and not keep_taps
        self._keepalive = []
        self.graph = None
        self._build()
        self._assign_buffers()
        for op in self.ops:
            op.fn = op.make()

    @property
    def input_images(self):
        """The Val the caller's images are copied into (uint8 for a u8_input plan)."""
        return getattr(self, "images_u8", None) or self.images

    # -------------------------------------------------------------- helpers
    def val(self, shape, dtype=None, name=None, keep=False):
        v = Val(shape, self.dtype if dtype is None else dtype, name, keep or self.keep_taps)
        self.vals.append(v)
        return v

    def add(self, kind, inputs, outputs, make, name="", flops=0, extra_bytes=0):
        ins = [i for i in inputs if i is not None]
        nbytes = sum(v.nbytes for v in ins) + sum(v.nbytes for v in outputs) + extra_bytes
        self.ops.append(Op(kind, ins, outputs, make, name, nbytes, flops))

    def w(self, key):
        return self.net.weights[key]

    def panel_for(self, key, taps, cin, cout, mode):
        """bf16 weight panel of the tensor-core path.  Inference: static (rebuilt by
        Network.refresh() when weights change)."""
        return self.net.static_panel(key, taps, cin, cout, mode)

    def folded(self, bn_name):
        return self.net.folded[bn_name]

    # -------------------------------------------------------------- op emitters
    def conv(self, xs, ys, weight_key, cin, cout, k=1, stride=1, scale=None, shift=None, act=ACT_NONE,
             gate=None, keep=None, residuals=None, ldc=None, ybs=None, y_offsets=None,
             in_dtype=None, out_dtype=None, name="", pre_split=False):
        """xs / ys: lists of Vals (one per group).  y_offsets: byte offsets into ys[i]."""
        n = len(xs)
        residuals = residuals or [None] * n
        wt = self.w(weight_key)
        in_dt = self.dtype if in_dtype is None else in_dtype
        out_dt = self.dtype if out_dtype is None else out_dtype
        # tcgen05 path: bf16 activations, channel count a multiple of 8 (TMA strides); stride 2
        # (BiFPN P6/P7 laterals) samples every other pixel through the tensor map's element strides
        use_tc = (self.net.use_tensor_cores and in_dt == BF16 and cin % 8 == 0 and
                  (stride == 1 or (stride == 2 and gate is None)))
        # fp32 accuracy mode of an INFERENCE plan: the same tensor-core kernel on the bf16 hi | lo split of the fp32
        # activations, three-term product accumulated in fp32 (effdet_conv_desc.split_planes); training plans keep
        # the exact SIMT kernel (their fp32 tests pin gradients to 1e-5)
        split = pre_split or (self.fp32_tensor_cores() and in_dt == F32 and out_dt == F32 and cin % 8 == 0 and
                              (stride == 1 or (stride == 2 and gate is None)))
        panel = gate_panel = None
        if split:
            lib = _lib.load()
            if not pre_split:            # pre_split: the producer already wrote the (B,H,W,2*cin) hi | lo planes
                xs_f32 = xs
                xs = [self.val(tuple(x.shape[:3]) + (2 * cin,), BF16, name + "_split%d" % i)
                      for i, x in enumerate(xs_f32)]
                for xf, xv in zip(xs_f32, xs):
                    rows = xf.shape[0] * xf.shape[1] * xf.shape[2]
                    self.add("split", [xf], [xv],
                             (lambda xf=xf, xv=xv, rows=rows: _call("effdet_split_bf16", xf.ptr, xv.ptr, rows, cin)),
                             name + "_split")
            if gate is not None:
                gate_panel = self.val((lib.effdet_conv_weight_panel_split_elems(self.B, cin, cout),), BF16,
                                      name + "_gated_panel")
                self.add("panel", [gate], [gate_panel],
                         lambda: _call("effdet_conv_weight_panel_split", wt.data_ptr(), gate_panel.ptr, 1, cin, cout,
                                       gate.ptr, self.B), name + "_panel")
            else:
                panel = self.net.static_panel(weight_key, k * k, cin, cout, "split")
            in_dt = BF16
            use_tc = True
        elif use_tc and gate is not None:
            lib = _lib.load()
            gate_panel = self.val((lib.effdet_conv_weight_panel_elems(self.B, cin, cout),), BF16,
                                  name + "_gated_panel")
            self.add("panel", [gate], [gate_panel],
                     lambda: _call("effdet_conv_weight_panel", wt.data_ptr(), gate_panel.ptr, 1, cin, cout,
                                   0, gate.ptr, self.B), name + "_panel")
        elif use_tc:
            panel = self.panel_for(weight_key, k * k, cin, cout, 0)

        def make():
            d = _lib.ConvDesc()
            d.n_groups = n
            for i in range(n):
                d.x[i] = xs[i].ptr
                d.y[i] = ys[i].ptr + (y_offsets[i] if y_offsets else 0)
                d.residual[i] = residuals[i].ptr if residuals[i] is not None else None
                d.H[i], d.W[i] = xs[i].shape[1], xs[i].shape[2]
                d.ldc[i] = ldc[i] if ldc else 0
                d.y_batch_stride[i] = ybs[i] if ybs else 0
            d.B, d.Cin, d.Cout, d.kh, d.kw, d.stride = self.B, cin, cout, k, k, stride
            d.weight = wt.data_ptr()
            d.scale = scale.data_ptr() if scale is not None else None
            d.shift = shift.data_ptr() if shift is not None else None
            d.gate = gate.ptr if gate is not None else None
            d.keep = keep.data_ptr() if keep is not None else None
            d.act, d.in_dtype, d.out_dtype = act, in_dt, out_dt
            d.split_planes = 1 if split else 0
            if gate_panel is not None:
                d.weight_bf16, d.weight_per_sample = gate_panel.ptr, 1
            elif panel is not None:
                d.weight_bf16, d.weight_per_sample = (panel.ptr if isinstance(panel, Val)
                                                      else panel.data_ptr()), 0
            else:
                d.weight_bf16 = None
            d.allow_tensor_core = 1 if use_tc else 0
            self._keepalive.append(d)
            return _call("effdet_conv2d", ctypes.byref(d))
        ins = list(xs) + [r for r in residuals if r is not None] + ([gate] if gate is not None else [])
        if gate_panel is not None:
            ins.append(gate_panel)
        if isinstance(panel, Val):
            ins.append(panel)
        out_es = 2 if out_dt == BF16 else 4
        flops = nbytes = 0
        for xv in xs:
            ho, wo = -(-xv.shape[1] // stride), -(-xv.shape[2] // stride)
            flops += 2 * self.B * ho * wo * cout * k * k * cin
            nbytes += self.B * ho * wo * cout * out_es
        nbytes += sum(v.nbytes for v in ins) + k * k * cin * cout * 4
        kind = "conv%dx%d" % (k, k) + ("_head" if n > 1 else "") + ("_tc" if use_tc else "")
        self.ops.append(Op(kind, ins, list(dict.fromkeys(ys)), make, name, nbytes, flops))

    # -------------------------------------------------------------- network lowering
    def _build(self):
        net, B, S = self.net, self.B, self.net.image_size
        bb = net.backbone
        self.images = self.val((B, S, S, 3), "u8" if self.u8_input else F32, "images", keep=True)
        x, H = self._stem()
        feats = []
        for bi, blk in enumerate(bb.blocks):
            x, H = self._mbconv(x, blk, H)
            if bi in bb.feature_after:
                x.keep = True
                feats.append(x)
                self.taps["C%d" % len(feats)] = x
        self.features = feats
        self._build_neck_and_heads(feats)

    def _stem(self):
        net, B, S = self.net, self.B, self.net.image_size
        H = (S + 1) // 2
        c0 = net.backbone.stem_filters
        x = self.val((B, H, H, c0), name="stem")
        sc, sh = self.folded("stem_bn")
        if self.u8_input:
            lut = net.normalization_lut()
            self.add("stem", [self.images], [x],
                     lambda: _call("effdet_stem_conv_u8", self.images.ptr, lut.data_ptr(),
                                   self.w("stem_conv/kernel").data_ptr(), sc.data_ptr(), sh.data_ptr(), x.ptr,
                                   B, S, S, c0, ACT_SWISH, self.dtype), "stem_conv")
        else:
            self.add("stem", [self.images], [x],
                     lambda: _call("effdet_stem_conv", self.images.ptr, self.w("stem_conv/kernel").data_ptr(),
                                   sc.data_ptr(), sh.data_ptr(), x.ptr, B, S, S, c0, self.dtype),
                     "stem_conv")
        return x, H

    def drop_keep(self, blk):
        """(B,) f32 scale of the projected branch of a skip block, or None.  Inference: FixedDropout is the
        identity (efficientnet.py:300-303 only acts in the training phase)."""
        return self.net.drop_scale.get(blk.prefix) if blk.has_skip else None

    def _mbconv(self, x, blk, H):
        """One MBConv block (efficientnet.py:210-306), inference form: BN folded into the conv /
        depthwise epilogues, SE gate folded into the project conv."""
        B = self.B
        lib = _lib.load()
        p = blk.prefix
        inp, cin, cmid, cout = x, blk.input_filters, blk.mid_filters, blk.output_filters
        if blk.expand_ratio != 1:
            e = self.val((B, H, H, cmid), name=p + "expand")
            s1, b1 = self.folded(p + "expand_bn")
            self.conv([x], [e], p + "expand_conv/kernel", cin, cmid, scale=s1, shift=b1,
                      act=ACT_SWISH, name=p + "expand_conv")
            x = e
        Ho = (H + blk.stride - 1) // blk.stride
        s2, b2 = self.folded(p + "bn")
        nblk = lib.effdet_dwconv_se_blocks(B, H, H, cmid, blk.stride, self.dtype)
        part = self.val((B, nblk, cmid), F32, name=p + "se_partial")
        # fp32 accuracy mode on the tensor cores: the depthwise output feeds only the project convolution, so the
        # kernel writes it as the bf16 hi | lo split that convolution reads (no fp32 copy, no separate split pass)
        dw_split = self.fp32_tensor_cores() and cmid % 8 == 0 and not self.keep_taps
        if dw_split:
            d = self.val((B, Ho, Ho, 2 * cmid), BF16, name=p + "dw_split")
            self.add("dwconv", [x], [d, part],
                     lambda: _call("effdet_dwconv_split_out", x.ptr, self.w(p + "dwconv/depthwise_kernel").data_ptr(),
                                   s2.data_ptr(), b2.data_ptr(), d.ptr, part.ptr, nblk, B, H, H, cmid,
                                   blk.kernel_size, blk.stride, ACT_SWISH), p + "dwconv",
                     flops=2 * blk.kernel_size ** 2 * B * Ho * Ho * cmid)
        else:
            d = self.val((B, Ho, Ho, cmid), name=p + "dw")
            self.add("dwconv", [x], [d, part],
                     lambda: _call("effdet_dwconv", x.ptr, self.w(p + "dwconv/depthwise_kernel").data_ptr(),
                                   s2.data_ptr(), b2.data_ptr(), d.ptr, part.ptr, nblk, B, H, H, cmid,
                                   blk.kernel_size, blk.stride, ACT_SWISH, self.dtype), p + "dwconv",
                     flops=2 * blk.kernel_size ** 2 * B * Ho * Ho * cmid)
        gate = self.val((B, cmid), F32, name=p + "gate")
        self.add("se", [part], [gate],
                 lambda: _call("effdet_se_gate", part.ptr, nblk, 1.0 / float(Ho * Ho),
                               self.w(p + "se_reduce/kernel").data_ptr(), self.w(p + "se_reduce/bias").data_ptr(),
                               self.w(p + "se_expand/kernel").data_ptr(), self.w(p + "se_expand/bias").data_ptr(),
                               gate.ptr, B, cmid, blk.se_filters), p + "se")
        y = self.val((B, Ho, Ho, cout), name=p + "out")
        s3, b3 = self.folded(p + "project_bn")
        keep = self.drop_keep(blk)
        self.conv([d], [y], p + "project_conv/kernel", cmid, cout, scale=s3, shift=b3,
                  gate=gate, keep=keep, residuals=[inp] if blk.has_skip else None,
                  name=p + "project_conv", pre_split=dw_split)
        return y, Ho

    def _build_neck_and_heads(self, feats):
        net = self.net
        Wd = net.w_bifpn
        for i in range(net.d_bifpn):
            feats = self._bifpn_layer(feats, i, Wd)
            for j, f in enumerate(feats):
                self.taps["BiFPN_%d_P%d" % (i, j + 3)] = f
        self.pyramid = feats
        # ---- heads
        self._heads(feats, Wd)

    def _conv_block(self, x, name, cin, cout, k=1, stride=1):
        B = self.B
        Ho = (x.shape[1] + stride - 1) // stride
        y = self.val((B, Ho, Ho, cout), name=name)
        sc, sh = self.folded(name + "_bn")
        self.conv([x], [y], name + "_conv/kernel", cin, cout, k=k, stride=stride, scale=sc, shift=sh,
                  act=ACT_RELU, name=name + "_conv")
        return y

    def _node(self, in0, mode0, in1, in2, fuse_name, dw_name, C):
        B, H = self.B, in1.shape[1]
        out = self.val((B, H, H, C), name=dw_name)
        sc, sh = self.folded(dw_name + "_bn")
        fw = self.w(fuse_name + "/" + fuse_name) if self.net.weighted_bifpn else None
        self.add("bifpn_node", [in0, in1, in2], [out],
                 lambda: _call("effdet_bifpn_node", in0.ptr, mode0, in1.ptr,
                               in2.ptr if in2 is not None else None,
                               fw.data_ptr() if fw is not None else None, 1e-4,
                               self.w(dw_name + "_dconv/depthwise_kernel").data_ptr(),
                               sc.data_ptr(), sh.data_ptr(), out.ptr, B, H, H, C, self.dtype),
                 dw_name, flops=(2 * 9 + 6) * B * H * H * C)
        return out

    def _bifpn_layer(self, feats, i, Wd):
        pre = "BiFPN_%d_" % i
        if i == 0:
            _, _, C3, C4, C5 = feats
            P3 = self._conv_block(C3, pre + "P3", C3.shape[3], Wd)
            P4 = self._conv_block(C4, pre + "P4", C4.shape[3], Wd)
            P5 = self._conv_block(C5, pre + "P5", C5.shape[3], Wd)
            P6 = self._conv_block(C5, pre + "P6", C5.shape[3], Wd, 3, 2)
            P7 = self._conv_block(P6, pre + "P7", Wd, Wd, 3, 2)
        else:
            P3, P4, P5, P6, P7 = [self._conv_block(f, pre + "P%d" % (3 + j), Wd, Wd)
                                  for j, f in enumerate(feats)]

        def fn(j):
            k = 8 * i + j
            return "w_bi_fpn_add" if k == 0 else "w_bi_fpn_add_%d" % k
        UP, DOWN = 1, 2
        P6_td = self._node(P7, UP, P6, None, fn(0), pre + "U_P6", Wd)
        P5_td = self._node(P6_td, UP, P5, None, fn(1), pre + "U_P5", Wd)
        P4_td = self._node(P5_td, UP, P4, None, fn(2), pre + "U_P4", Wd)
        P3_o = self._node(P4_td, UP, P3, None, fn(3), pre + "U_P3", Wd)
        P4_o = self._node(P3_o, DOWN, P4_td, P4, fn(4), pre + "D_P4", Wd)
        P5_o = self._node(P4_o, DOWN, P5_td, P5, fn(5), pre + "D_P5", Wd)
        P6_o = self._node(P5_o, DOWN, P6_td, P6, fn(6), pre + "D_P6", Wd)
        P7_o = self._node(P6_o, DOWN, P7, None, fn(7), pre + "D_P7", Wd)
        return [P3_o, P4_o, P5_o, P6_o, P7_o]

    def _heads(self, feats, Wd):
        net, B = self.net, self.B
        C, A = net.num_classes, 9
        hw = [f.shape[1] * f.shape[2] for f in feats]
        N = A * sum(hw)
        self.N = N
        self.regression = self.val((B, N, 4), F32, "regression", keep=True)
        self.classification = self.val((B, N, C), F32, "classification", keep=True)
        lvl_off = np.concatenate([[0], np.cumsum([A * h for h in hw])[:-1]])
        for scope, fmt, final, out, per, act in (
                ("box_head", "regress_head_conv_%d", "regress_head_conv_final", self.regression, 4,
                 ACT_NONE),
                ("class_head", "class_head_%d", "pyramid_classification", self.classification, C,
                 ACT_SIGMOID)):
            xs = list(feats)
            for i in range(net.head_depth):
                ys = [self.val(x.shape, name="%s_%d_l%d" % (scope, i, l)) for l, x in enumerate(xs)]
                n = scope + "/" + fmt % i
                self.conv(xs, ys, n + "/kernel", Wd, Wd, k=3, shift=self.w(n + "/bias"), act=ACT_RELU,
                          name=n)
                xs = ys
            n = scope + "/" + final
            self.conv(xs, [out] * len(xs), n + "/kernel", Wd, A * per, k=3, shift=self.w(n + "/bias"),
                      act=act, ldc=[A * per] * len(xs), ybs=[N * per] * len(xs),
                      y_offsets=[int(o) * per * 4 for o in lvl_off], out_dtype=F32, name=n)

    # -------------------------------------------------------------- buffers
    def _assign_buffers(self):
        for idx, op in enumerate(self.ops):
            for v in op.inputs + op.outputs:
                v.last_use = max(v.last_use, idx)
        free = {}
        total = 0

        def alloc(v):
            nonlocal total
            if self.reuse and not v.keep:
                lst = free.get(v.nbytes)
                if lst:
                    v.t = lst.pop()
                    return
            v.t = torch.empty(max(v.nbytes, 16), dtype=torch.uint8, device=self.dev)
            total += v.nbytes
        done = set()
        for v in self.vals:
            if v.last_use < 0:            # e.g. images (only read) -- allocate anyway
                alloc(v); done.add(id(v))
        for idx, op in enumerate(self.ops):
            for v in op.outputs:
                if id(v) not in done:
                    alloc(v); done.add(id(v))
            for v in op.inputs:
                if id(v) not in done:
                    alloc(v); done.add(id(v))
            if self.reuse:
                for v in set(op.inputs + op.outputs):
                    if v.last_use == idx and not v.keep and v.t is not None:
                        free.setdefault(v.nbytes, []).append(v.t)
        self.activation_bytes = total

    def tensor(self, v):
        """torch view of a planned value (for tests / outputs)."""
        return v.t[:v.nbytes].view(TORCH_DTYPE[v.dtype]).view(v.shape)

    # -------------------------------------------------------------- execution
    def run(self, begin=0, end=None):
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        for op in self.ops[begin:end]:
            op.fn(stream)

    def capture(self, bounds=None):
        """Capture the launch list in a CUDA graph (replay with .replay()).  bounds: ascending launch indices
        [e0, e1, ..., len(ops)]: one graph per segment [0,e0), [e0,e1), ... (replay_segment(i)) so that the
        caller can interleave work -- the per-bucket gradient all-reduce -- between them."""
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self.run()              # warm-up outside capture
        torch.cuda.current_stream(self.dev).wait_stream(s)
        bounds = list(bounds) if bounds else [len(self.ops)]
        assert bounds[-1] == len(self.ops) and all(a < b for a, b in zip(bounds, bounds[1:]))
        self.segment_graphs = []
        lanes = self.graph_lanes()
        begin = 0
        import gc
        for end in bounds:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                # no cyclic garbage collection inside the capture: collecting an older plan (CUDA graphs, streams,
                # device buffers in reference cycles) calls cudaFree / cudaGraphExecDestroy, which invalidates a
                # capture in progress (run_lanes allocates enough Python objects to trigger a collection)
                was_on = gc.isenabled()
                gc.disable()
                try:
                    if lanes > 1:
                        self.run_lanes(begin, end, lanes)
                    else:
                        self.run(begin, end)
                finally:
                    if was_on:
                        gc.enable()
            self.segment_graphs.append(g)
            begin = end
        self.graph = self.segment_graphs[0] if len(self.segment_graphs) == 1 else self.segment_graphs
        return self.graph

    def graph_lanes(self):
        """Streams the launch list is captured on.  1 = the plan order as one chain.  More: independent chains
        (the two heads, the BiFPN lateral convolutions, weight-panel preparation) become parallel branches of the
        CUDA graph -- what a batch-1 step, a chain of ~120 launches of 3-9 us, is short of is not bandwidth but
        things to run side by side.  Default: 4 for small inference batches, EFFDET_GRAPH_LANES overrides."""
        env = os.environ.get("EFFDET_GRAPH_LANES")
        if env:
            return max(1, int(env))
        return 4 if (type(self) is Plan and self.B <= 2) else 1

    def lane_hint(self, op, n_lanes):
        return None

    def fp32_tensor_cores(self):
        """fp32 accuracy mode: run the dense convolutions on the tensor cores through the split-bf16 form
        (inference plans; EFFDET_FP32_TC=0 keeps the exact SIMT kernel)."""
        return (type(self) is Plan and self.net.use_tensor_cores and self.dtype == F32 and
                os.environ.get("EFFDET_FP32_TC", "1") != "0")

    def dependencies(self):
        """Per launch: the earlier launches it must follow -- last writer of every buffer it reads (RAW), last
        writer and the readers since of every buffer it writes (WAW / WAR; buffers are recycled between values,
        so hazards are tracked per BUFFER, not per value).  Launches exchange data only through their declared
        inputs / outputs; weights and static panels are read-only inside a plan; per-layer state outside the values
        (BatchNorm moving statistics, parameter gradients) has exactly one writer per step, or writers that
        lane_hint() keeps on one lane.  A launch that declares neither inputs nor outputs is a barrier."""
        last_w, readers, deps = {}, {}, []
        barrier = None
        for i, op in enumerate(self.ops):
            d = set()
            if barrier is not None:
                d.add(barrier)
            if not op.inputs and not op.outputs:     # declares nothing: pure side effect (the stochastic-depth
                d.update(range(0 if barrier is None else barrier, i))     # mask draw) -> ordered against everything
                barrier = i
            rb = {v.t.data_ptr() for v in op.inputs if v.t is not None}
            wb = {v.t.data_ptr() for v in op.outputs if v.t is not None}
            for b in rb:
                if b in last_w:
                    d.add(last_w[b])
            for b in wb:
                if b in last_w:
                    d.add(last_w[b])
                d.update(readers.get(b, ()))
            d.discard(i)
            deps.append(sorted(d))
            for b in rb - wb:
                readers.setdefault(b, set()).add(i)
            for b in wb:
                last_w[b] = i
                readers[b] = set()
        return deps

    def lane_schedule(self, begin, end, n_lanes):
        """[(lane, [launch indices on OTHER lanes to wait for])] for launches [begin, end): the multi-lane order
        run_lanes() enqueues and plan_export.py writes into a compiled plan.  A launch continues the lane of its
        latest dependency when that launch is the lane's tail, else takes an unused lane, else the lane idle
        longest; lane_hint() can pin it (training plans) or restrict the choice."""
        deps = self.dependencies()
        lane_last = [None] * n_lanes                     # last launch index enqueued on the lane
        lane_of, out = {}, []
        for i in range(begin, end):
            d = [j for j in deps[i] if j >= begin]
            hint = self.lane_hint(self.ops[i], n_lanes)
            lane = hint if isinstance(hint, int) else None
            allowed = list(hint) if isinstance(hint, (list, tuple)) else list(range(n_lanes))
            if lane is None:
                for j in sorted(d, reverse=True):
                    if lane_of[j] in allowed and lane_last[lane_of[j]] == j:
                        lane = lane_of[j]
                        break
            if lane is None:
                unused = [l for l in allowed if lane_last[l] is None]
                lane = unused[0] if unused else min(allowed, key=lambda l: lane_last[l])
            out.append((lane, [j for j in d if lane_of[j] != lane]))
            lane_of[i], lane_last[lane] = lane, i
        return out

    def run_lanes(self, begin, end, n_lanes):
        """Enqueue launches [begin, end) on up to n_lanes streams, joined by events according to lane_schedule().
        Called inside a stream capture (the fork / join events become graph edges); launch order within a lane is
        plan order."""
        origin = torch.cuda.current_stream(self.dev)
        sched = self.lane_schedule(begin, end, n_lanes)
        if not hasattr(self, "_lane_streams") or len(self._lane_streams) < n_lanes - 1:
            self._lane_streams = [torch.cuda.Stream(self.dev) for _ in range(n_lanes - 1)]
        streams = [origin] + self._lane_streams[:n_lanes - 1]
        start = torch.cuda.Event()
        start.record(origin)
        joined = [True] + [False] * (n_lanes - 1)        # lane has forked from the capturing stream
        lane_tail = [None] * n_lanes
        done_ev = {}
        for i, (lane, waits) in zip(range(begin, end), sched):
            st = streams[lane]
            if not joined[lane]:
                st.wait_event(start)
                joined[lane] = True
            for j in waits:
                st.wait_event(done_ev[j])
            self.ops[i].fn(st.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(st)
            done_ev[i], lane_tail[lane] = ev, i
        for l in range(1, n_lanes):                      # join: the capture ends on the origin stream
            if lane_tail[l] is not None:
                origin.wait_event(done_ev[lane_tail[l]])

    def replay(self):
        if self.graph is None:
            self.run()
        else:
            for g in self.segment_graphs:
                g.replay()

    def replay_segment(self, i):
        self.segment_graphs[i].replay()

    def profile(self, iters=3, repeat=8):
        """Per-launch device times -> list of dicts {name, kind, ms, bytes, flops}.
        Every launch is captured `repeat` times back to back into its own small CUDA graph and the
        graph is replayed between two CUDA events on the launching stream: the interval holds only
        device time (kernel + the inter-node gap it also has inside the plan's graph), not the host
        work of building descriptors / tensor maps, which an eager event pair around a 3-microsecond
        kernel would mostly measure.  Large launches stream more bytes than L2 holds; small ones
        are L2-resident here as they are in the real step (their producer just wrote their input)."""
        stream = torch.cuda.current_stream(self.dev)
        # EFFDET_PROFILE_KINDS=dwconv,conv1x1_tc: the eager warm-up pass brackets the launches of those kinds with
        # cudaProfilerStart / Stop, so `ncu --profile-from-start off` captures exactly one launch per op, in plan
        # order (profiles/tools/traffic_from_ncu.py turns that capture into profiles/traffic.json)
        kinds = set(filter(None, os.environ.get("EFFDET_PROFILE_KINDS", "").split(",")))
        for op in self.ops:                      # warm (kernel attributes, lazy allocations)
            if op.kind in kinds:
                torch.cuda.synchronize(self.dev)
                torch.cuda.profiler.start()
            op.fn(stream.cuda_stream)
            if op.kind in kinds:
                torch.cuda.synchronize(self.dev)
                torch.cuda.profiler.stop()
        torch.cuda.synchronize(self.dev)
        graphs = []
        side = torch.cuda.Stream(self.dev)
        for op in self.ops:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                sp = torch.cuda.current_stream(self.dev).cuda_stream
                for _ in range(repeat):
                    op.fn(sp)
            graphs.append(g)
        acc = [0.0] * len(self.ops)
        for g in graphs:
            g.replay()
        torch.cuda.synchronize(self.dev)
        for _ in range(iters):
            evs = []
            for g in graphs:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                g.replay()
                b.record(stream)
                evs.append((a, b))
            torch.cuda.synchronize(self.dev)
            for i, (a, b) in enumerate(evs):
                acc[i] += a.elapsed_time(b)
        return [dict(name=op.name, kind=op.kind, ms=acc[i] / (iters * repeat), bytes=op.nbytes, flops=op.flops)
                for i, op in enumerate(self.ops)]

    def forward(self, images):
        """images: (B,S,S,3) float32 (or, for a u8_input plan, uint8) CUDA tensor -> (regression,
        classification) views."""
        self.tensor(self.input_images).copy_(images, non_blocking=True)
        self.replay()
        return self.tensor(self.regression), self.tensor(self.classification)
