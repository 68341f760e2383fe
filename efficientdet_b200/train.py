"""Training step on one GPU (+ data-parallel all-reduce across ranks).

Replaces what Keras/TF did inside `model.fit` for the reference (train_tpu.py:249-346, call stack
SURVEY section 3 (C)): forward in training mode (BatchNorm batch statistics in every layer that
is trainable and not freeze_bn), focal + smooth-L1 losses, backward through heads and BiFPN,
gradient all-reduce (tf.distribute.MirroredStrategy -> NCCL over NVLink here), SGD-momentum.

Scope of this build: layers up to the backbone boundary must be frozen
(`--freeze-backbone`, train_tpu.py:272-274: model.layers[1:EFFICIENTNET_DEPTHS[phi]]); the
backbone then runs in inference mode and receives no gradient.  Full-backbone backward is listed
under "next" in DESIGN.md.
"""
import ctypes
import os
import time

import numpy as np
import torch

from . import _lib, engine
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SWISH, BF16, F32
from .engine import (BN_EPS_BACKBONE, BN_EPS_BIFPN, BN_MOMENTUM_BACKBONE, BN_MOMENTUM_BIFPN, Op, Val,
                     _call)

UP, DOWN = 1, 2


class TrainPlan(engine.Plan):
    """Forward (training mode) + losses + backward launch list for a fixed batch."""

    def __init__(self, net, batch, alpha=0.25, gamma=1.5, delta=1.0, dense_labels=False,
                 reuse_buffers=True, train_backbone=False, u8_input=False):
        self.alpha, self.gamma, self.delta = alpha, gamma, delta
        self.dense_labels = dense_labels
        self.train_backbone = train_backbone
        self.tape = []
        self.bb_tape = []
        self.gvals = {}            # id(Val) -> gradient Val
        net.ensure_grad_buffers()
        # stochastic depth (efficientnet.py:300-304): one (B,) scale vector per skip block with drop_rate > 0,
        # redrawn on the device at the start of every step
        self.drop_blocks = [b for b in net.backbone.blocks if b.has_skip and b.drop_rate and b.drop_rate > 0]
        self.drop_index = {b.prefix: i for i, b in enumerate(self.drop_blocks)}
        if self.drop_blocks:
            dev = net.device
            self.drop_rates = torch.tensor([float(b.drop_rate) for b in self.drop_blocks], dtype=torch.float32,
                                           device=dev)
            self.drop_scales = torch.ones((len(self.drop_blocks), int(batch)), dtype=torch.float32, device=dev)
            self.drop_step = torch.zeros((1,), dtype=torch.int64, device=dev)
            rank = torch.distributed.get_rank() if (torch.distributed.is_available() and
                                                    torch.distributed.is_initialized()) else 0
            self.drop_seed = (int(getattr(net, "seed", 0) or 0) * 1000003 + 7919 * rank + 12345) & (2 ** 63 - 1)
        super().__init__(net, batch, reuse_buffers=reuse_buffers, u8_input=u8_input)

    def capture(self):
        """Capturing warms the launch list up with one real run (engine.Plan.capture): that run has side effects
        on a training plan -- it advances the BatchNorm moving averages and the stochastic-depth step counter --
        so both are snapshotted and restored around it (the captured graph itself does not execute)."""
        net = self.net
        snap = net.flat.clone()
        step = self.drop_step.clone() if self.drop_blocks else None
        from . import parallel
        self.buckets = None
        if parallel.world()[1] > 1 or os.environ.get("EFFDET_FORCE_BUCKETS") == "1":
            min_elems = int(os.environ.get("EFFDET_BUCKET_MIN_ELEMS", 1 << 20))
            self.buckets = parallel.plan_buckets(self.bucket_marks, net.flat.numel(), min_elems)
        g = super().capture([b[0] for b in self.buckets] if self.buckets else None)
        torch.cuda.current_stream(self.dev).synchronize()
        net.flat.copy_(snap)
        if step is not None:
            self.drop_step.copy_(step)
        return g

    # Weight gradients are the leaves of the backward pass: nothing in the step reads them before the optimizer, and
    # they write only their own scratch partials and the (untracked) gradient buffers.  Captured on a second lane
    # they run beside the data-gradient chain, whose BiFPN / head kernels at the coarse pyramid levels are a few
    # microseconds on a handful of SMs (engine.Plan.run_lanes; hazards on recycled buffers are tracked per buffer).
    WGRAD_KINDS = frozenset(("conv_wgrad_tc", "conv_wgrad", "dw_wgrad", "fuse_wgrad", "bias_grad", "stem_wgrad"))

    # Operand preparation that depends on the weights alone (bf16 panels of the trainable convolutions, transposed /
    # flipped kernels for the data gradients): ready when the step starts, so a third lane runs it beside the
    # backbone forward instead of in front of each consumer.
    PREP_KINDS = frozenset(("panel", "wtrans"))

    def graph_lanes(self):
        env = os.environ.get("EFFDET_GRAPH_LANES")
        return max(1, int(env)) if env else 4

    def lane_hint(self, op, n_lanes):
        if op.kind in self.WGRAD_KINDS and n_lanes > 1:
            return 1
        if op.kind in self.PREP_KINDS and not op.inputs and n_lanes > 2:
            return 2
        if n_lanes > 3:                       # the data path itself over lanes 0, 3, 4, ... by dependencies
            return [0] + list(range(3, n_lanes))
        return 0

    # ------------------------------------------------------------------ small helpers
    def gw(self, key):
        return self.net.grads[key]

    def fvec(self, C, name):
        return self.val((C,), F32, name, keep=True)

    def _scratch(self, nfloats, name):
        return self.val((int(nfloats),), F32, name)

    def bn_is_training(self, bn_name):
        return not (self.net.freeze_bn or bn_name in self.net.frozen_layers)

    def _bn_train(self, z, y, bn_name, C, rows, act, dw_from=None):
        """z -> y = act(BN_batch(z)); returns the record needed by the backward pass.
        dw_from = (f, kernel_key, H): z is the raw 3x3 stride-1 depthwise conv of f, not computed yet --
        in bf16 mode with a trainable BN the depthwise kernel then also emits the batch statistics."""
        lib = _lib.load()
        is_bifpn = bn_name.startswith("BiFPN_")
        eps = BN_EPS_BIFPN if is_bifpn else BN_EPS_BACKBONE
        mom = BN_MOMENTUM_BIFPN if is_bifpn else BN_MOMENTUM_BACKBONE
        rec = dict(bn=bn_name, C=C, rows=rows, train=self.bn_is_training(bn_name), act=act)
        if dw_from is not None:
            f, kkey, H = dw_from[:3]
            dk, ds = (dw_from[3], dw_from[4]) if len(dw_from) > 3 else (3, 1)
            dHo = -(-H // ds)
            B = self.B
            ones, zeros = self.net.const_ones(C), self.net.const_zeros(C)
            if rec["train"] and self.dtype == BF16:
                nblk = B * lib.effdet_dwconv_se_blocks(B, H, H, C, ds, self.dtype)
                sc, sh = self.fvec(C, bn_name + "/scale_t"), self.fvec(C, bn_name + "/shift_t")
                mu, iv = self.fvec(C, bn_name + "/mean_t"), self.fvec(C, bn_name + "/invstd_t")
                part = self._scratch(2 * C * nblk, bn_name + "/partial")
                w = self.w
                self.add("dwconv", [f], [z, sc, sh, mu, iv, part],
                         lambda: _call("effdet_dwconv_bn_stats", f.ptr, w(kkey).data_ptr(), ones.data_ptr(),
                                       zeros.data_ptr(), z.ptr, B, H, H, C, dk, ds, w(bn_name + "/gamma").data_ptr(),
                                       w(bn_name + "/beta").data_ptr(), eps, mom,
                                       w(bn_name + "/moving_mean").data_ptr(),
                                       w(bn_name + "/moving_variance").data_ptr(), sc.ptr, sh.ptr, mu.ptr, iv.ptr,
                                       part.ptr, nblk, self.dtype), bn_name[:-3] + "_dconv+stats",
                         flops=2 * dk * dk * B * dHo * dHo * C)
                self.add("bn_apply", [z, sc, sh], [y],
                         lambda: _call("effdet_scale_shift_act", z.ptr, sc.ptr, sh.ptr, y.ptr, rows, C, act,
                                       self.dtype), bn_name + "_apply")
                rec.update(mean=mu, invstd=iv, nblk=lib.effdet_colreduce_blocks(rows, C, self.dtype), ua=sc, ub=sh)
                return rec
            self.add("dwconv", [f], [z],
                     lambda: _call("effdet_dwconv", f.ptr, self.w(kkey).data_ptr(), ones.data_ptr(),
                                   zeros.data_ptr(), z.ptr, None, 0, B, H, H, C, dk, ds, ACT_NONE, self.dtype),
                     bn_name[:-3] + "_dconv", flops=2 * dk * dk * B * dHo * dHo * C)
        if rec["train"]:
            nblk = lib.effdet_colreduce_blocks(rows, C, self.dtype)
            sc, sh = self.fvec(C, bn_name + "/scale_t"), self.fvec(C, bn_name + "/shift_t")
            mu, iv = self.fvec(C, bn_name + "/mean_t"), self.fvec(C, bn_name + "/invstd_t")
            part = self._scratch(2 * C * nblk, bn_name + "/partial")
            w = self.w
            self.add("bn_stats", [z], [sc, sh, mu, iv, part],
                     lambda: _call("effdet_bn_train_stats", z.ptr, rows, C, w(bn_name + "/gamma").data_ptr(),
                                   w(bn_name + "/beta").data_ptr(), eps, mom,
                                   w(bn_name + "/moving_mean").data_ptr(),
                                   w(bn_name + "/moving_variance").data_ptr(), sc.ptr, sh.ptr, mu.ptr,
                                   iv.ptr, part.ptr, nblk, self.dtype), bn_name + "_stats")
            self.add("bn_apply", [z, sc, sh], [y],
                     lambda: _call("effdet_scale_shift_act", z.ptr, sc.ptr, sh.ptr, y.ptr, rows, C, act,
                                   self.dtype), bn_name + "_apply")
            rec.update(mean=mu, invstd=iv, nblk=nblk, ua=sc, ub=sh)
        else:
            fs, fb = self.folded(bn_name)
            self.add("bn_apply", [z], [y],
                     lambda: _call("effdet_scale_shift_act", z.ptr, fs.data_ptr(), fb.data_ptr(), y.ptr,
                                   rows, C, act, self.dtype), bn_name + "_apply")
        return rec

    def panel_for(self, key, taps, cin, cout, mode):
        """Trainable weights change every step: their bf16 panels are rebuilt inside the step;
        frozen-backbone panels stay static."""
        net = self.net
        if net.offsets[key] < net.backbone_end:
            return net.static_panel(key, taps, cin, cout, mode)
        ck = (key, mode)
        if not hasattr(self, "_step_panels"):
            self._step_panels = {}
        if ck not in self._step_panels:
            n = _lib.load().effdet_conv_weight_panel_elems(taps, cin if mode == 0 else cout,
                                                           cout if mode == 0 else cin)
            pv = self.val((n,), BF16, key + "_panel%d" % mode, keep=True)
            self.add("panel", [], [pv],
                     lambda: _call("effdet_conv_weight_panel", self.w(key).data_ptr(), pv.ptr, taps, cin,
                                   cout, mode, None, 0), key + "_panel%d" % mode)
            self._step_panels[ck] = pv
        return self._step_panels[ck]

    # ------------------------------------------------------------------ forward emitters (override)
    def _conv_block(self, x, name, cin, cout, k=1, stride=1):
        B = self.B
        Ho = (x.shape[1] + stride - 1) // stride
        z = self.val((B, Ho, Ho, cout), name=name + "_z", keep=True)
        y = self.val((B, Ho, Ho, cout), name=name, keep=True)
        self.conv([x], [z], name + "_conv/kernel", cin, cout, k=k, stride=stride, name=name + "_conv")
        rec = self._bn_train(z, y, name + "_bn", cout, B * Ho * Ho, ACT_RELU)
        rec.update(kind="convblock", x=x, z=z, y=y, name=name, cin=cin, cout=cout, k=k, stride=stride)
        self.tape.append(rec)
        return y

    def _node(self, in0, mode0, in1, in2, fuse_name, dw_name, C):
        B, H = self.B, in1.shape[1]
        net = self.net
        f = self.val((B, H, H, C), name=dw_name + "_f", keep=True)
        z = self.val((B, H, H, C), name=dw_name + "_z", keep=True)
        y = self.val((B, H, H, C), name=dw_name, keep=True)
        fw = self.w(fuse_name + "/" + fuse_name) if net.weighted_bifpn else None
        fwp = fw.data_ptr() if fw is not None else None
        self.add("fuse", [in0, in1, in2], [f],
                 lambda: _call("effdet_resample_fuse", in0.ptr, mode0, in1.ptr,
                               in2.ptr if in2 is not None else None, fwp, 1e-4, f.ptr, B, H, H, C,
                               self.dtype), dw_name + "_fuse")
        rec = self._bn_train(z, y, dw_name + "_bn", C, B * H * H, ACT_RELU,
                             dw_from=(f, dw_name + "_dconv/depthwise_kernel", H))
        rec.update(kind="node", in0=in0, mode0=mode0, in1=in1, in2=in2, f=f, z=z, y=y, fuse=fuse_name,
                   name=dw_name, H=H)
        self.tape.append(rec)
        return y

    def _heads(self, feats, Wd):
        start = len(self.ops)
        super()._heads(feats, Wd)
        # recover the per-layer activations from the conv ops just emitted
        net = self.net
        n_ops = 2 * (net.head_depth + 1)
        ops = [op for op in self.ops[start:] if op.kind.startswith("conv")]
        assert len(ops) == n_ops
        self.head_tape = []
        for h, (scope, fmt, final, out, per) in enumerate((
                ("box_head", "regress_head_conv_%d", "regress_head_conv_final", self.regression, 4),
                ("class_head", "class_head_%d", "pyramid_classification", self.classification,
                 net.num_classes))):
            layers = []
            xs = list(feats)
            for i in range(net.head_depth):
                op = ops[h * (net.head_depth + 1) + i]
                ys = op.outputs
                for v in ys:
                    v.keep = True
                layers.append(dict(name=scope + "/" + fmt % i, xs=xs, ys=ys))
                xs = ys
            self.head_tape.append(dict(scope=scope, final=scope + "/" + final, layers=layers, xs=xs,
                                       out=out, per=per))
        for f in feats:
            f.keep = True

    # ------------------------------------------------------------------ whole step
    def _build(self):
        super()._build()                      # forward (training-mode emitters above)
        self.n_forward_ops = len(self.ops)
        self._build_losses()
        self._build_backward()

    def _build_losses(self):
        B, N, C = self.B, self.N, self.net.num_classes
        self.reg_t = self.val((B, N, 5), F32, "regression_targets", keep=True)
        if self.dense_labels:
            self.lab_t = self.val((B, N, C + 1), F32, "label_targets", keep=True)
            self.state_t = self.cls_t = None
        else:
            self.lab_t = None
            self.state_t = self.val((B, N), "i8", "state_targets", keep=True)
            self.cls_t = self.val((B, N), "i32", "class_targets", keep=True)
        self.dcls = self.val((B, N, C), F32, "dclassification_logits", keep=True)
        self.dreg = self.val((B, N, 4), F32, "dregression", keep=True)
        self.loss_out = self.val((8,), F32, "losses", keep=True)
        ws_bytes = _lib.load().effdet_detection_losses_workspace_size()
        ws = self._scratch(ws_bytes // 4, "loss_ws")
        ins = [self.classification, self.regression, self.reg_t, self.lab_t, self.state_t, self.cls_t]
        # tensor-core gradient kernels read bf16, channel-padded, per-level copies of dreg / dcls
        self.use_tc_grads = self.net.use_tensor_cores and self.dtype == BF16
        self.lvl_dcls = self.lvl_dreg = None
        outs = [self.dcls, self.dreg, self.loss_out, ws]
        if self.use_tc_grads:
            feats = self.pyramid
            self.cpad_cls, self.cpad_reg = -(-9 * C // 8) * 8, 40
            self.lvl_dcls = [self.val((B, f.shape[1], f.shape[2], self.cpad_cls), BF16, "dcls_l%d" % i, keep=True)
                             for i, f in enumerate(feats)]
            self.lvl_dreg = [self.val((B, f.shape[1], f.shape[2], self.cpad_reg), BF16, "dreg_l%d" % i, keep=True)
                             for i, f in enumerate(feats)]
            outs += self.lvl_dcls + self.lvl_dreg
            cells = (ctypes.c_int * 5)(*[f.shape[1] * f.shape[2] for f in feats])
        # the concatenated fp32 gradients have one reader left in tensor-core mode: the bias gradient of the two
        # final head convolutions, and only when the weight-gradient kernel does not produce it (Cin > 64)
        self.fp32_head_grads = not (self.use_tc_grads and self.wgrad_fuses_bias(self.net.w_bifpn, 3))

        def make():
            args = [self.classification.ptr, self.regression.ptr,
                    self.reg_t.ptr, self.lab_t.ptr if self.lab_t is not None else None,
                    self.state_t.ptr if self.state_t is not None else None,
                    self.cls_t.ptr if self.cls_t is not None else None, B, N, C, self.alpha,
                    self.gamma, self.delta, 1.0, self.dcls.ptr if self.fp32_head_grads else None,
                    self.dreg.ptr if self.fp32_head_grads else None, self.loss_out.ptr, ws.ptr, ws_bytes]
            if self.use_tc_grads:
                for v in self.lvl_dcls + self.lvl_dreg:      # zero the padding channels once
                    self.tensor(v).zero_()
                pc = (ctypes.c_void_p * 5)(*[v.ptr for v in self.lvl_dcls])
                pr = (ctypes.c_void_p * 5)(*[v.ptr for v in self.lvl_dreg])
                self._keepalive += [pc, pr, cells]
                args += [pc, pr, cells, 5, self.cpad_cls, self.cpad_reg]
            else:
                args += [None, None, None, 0, 0, 0]
            return _call("effdet_detection_losses", *args)
        self.add("losses", ins, outs, make, "losses")

    # gradient bookkeeping ------------------------------------------------------------
    def grad_of(self, v):
        """(gradient Val, accumulate?) for activation v; first writer overwrites."""
        g = self.gvals.get(id(v))
        if g is None:
            g = self.val(v.shape, v.dtype, (v.name or "") + "_grad")
            self.gvals[id(v)] = g
            return g, 0
        return g, 1

    def _transposed_weight(self, key, taps, cin, cout):
        wt = self.val((taps * cin * cout,), F32, key + "_T")
        self.add("wtrans", [], [wt],
                 lambda: _call("effdet_conv_weight_transpose", self.w(key).data_ptr(), wt.ptr, taps, cin,
                               cout), key + "_T")
        return wt

    def wgrad_fuses_bias(self, cin, k):
        """True when the tensor-core weight gradient of a (k x k, cin -> *) convolution also produces the
        bias gradient (effdet_conv_wgrad_tc_fuses_bias)."""
        return bool(self.net.use_tensor_cores and self.dtype == BF16 and cin <= 64 and cin % 8 == 0 and k == 3)

    def _wgrad_tc(self, xs, dzs, key, cin, cout, k, dz_ld=None, name="", bias_key=None):
        """bf16 operands, stride 1: tcgen05 weight gradient (+ the bias gradient when bias_key is given
        and the kernel can fuse it)."""
        lib = _lib.load()
        n = len(xs)
        d = _lib.WgradDesc()
        d.n_groups = n
        for i in range(n):
            d.H[i], d.W[i] = xs[i].shape[1], xs[i].shape[2]
            d.dz_ld[i] = dz_ld if dz_ld else 0
        d.B, d.Cin, d.Cout, d.kh, d.kw, d.stride = self.B, cin, cout, k, k, 1
        nsplit = lib.effdet_conv_wgrad_tc_splits(ctypes.byref(d))
        if bias_key is not None:
            assert lib.effdet_conv_wgrad_tc_fuses_bias(ctypes.byref(d))
        part = self._scratch(nsplit * (k * k * cin * cout + (cout if bias_key is not None else 0)),
                             key + "_wg_partial")
        gw = self.gw(key)
        gb = self.gw(bias_key) if bias_key is not None else None

        def make():
            for i in range(n):
                d.x[i], d.dz[i] = xs[i].ptr, dzs[i].ptr
            d.dweight, d.partial, d.n_splits, d.accumulate = gw.data_ptr(), part.ptr, nsplit, 0
            d.dbias = gb.data_ptr() if gb is not None else None
            d.x_dtype = d.dz_dtype = BF16
            self._keepalive.append(d)
            return _call("effdet_conv_wgrad_tc", ctypes.byref(d))
        flops = sum(2 * self.B * x.shape[1] * x.shape[2] * cin * cout * k * k for x in xs)
        self.ops.append(Op("conv_wgrad_tc", list(xs) + list(dzs), [part], make, name,
                           sum(v.nbytes for v in xs) + sum(v.nbytes for v in dzs), flops))

    def _wgrad(self, xs, dzs, key, cin, cout, k, stride, dz_ld=None, dz_bs=None, dz_off=None,
               dz_dtype=None, name="", bias_key=None):
        if (self.net.use_tensor_cores and self.dtype == BF16 and stride == 1 and cin % 8 == 0
                and dz_off is None and (dz_dtype is None or dz_dtype == BF16)):
            return self._wgrad_tc(xs, dzs, key, cin, cout, k, name=name, bias_key=bias_key)
        assert bias_key is None
        lib = _lib.load()
        n = len(xs)
        d = _lib.WgradDesc()
        d.n_groups = n
        for i in range(n):
            d.H[i], d.W[i] = xs[i].shape[1], xs[i].shape[2]
        d.B, d.Cin, d.Cout, d.kh, d.kw, d.stride = self.B, cin, cout, k, k, stride
        nsplit = lib.effdet_conv_wgrad_splits(ctypes.byref(d))
        part = self._scratch(nsplit * k * k * cin * cout, key + "_wg_partial")
        gw = self.gw(key)
        xdt = self.dtype
        zdt = self.dtype if dz_dtype is None else dz_dtype

        def make():
            for i in range(n):
                d.x[i] = xs[i].ptr
                d.dz[i] = dzs[i].ptr + (dz_off[i] if dz_off else 0)
                d.dz_ld[i] = dz_ld[i] if dz_ld else 0
                d.dz_batch_stride[i] = dz_bs[i] if dz_bs else 0
            d.dweight, d.partial, d.n_splits, d.accumulate = gw.data_ptr(), part.ptr, nsplit, 0
            d.x_dtype, d.dz_dtype = xdt, zdt
            self._keepalive.append(d)
            return _call("effdet_conv_wgrad", ctypes.byref(d))
        flops = sum(2 * self.B * (-(-x.shape[1] // stride)) * (-(-x.shape[2] // stride)) * cin * cout * k * k
                    for x in xs)
        self.ops.append(Op("conv_wgrad", list(xs) + list(dict.fromkeys(dzs)), [part], make, name,
                           sum(v.nbytes for v in xs) + sum(v.nbytes for v in dict.fromkeys(dzs)), flops))

    def _bias_grad(self, dz, rows, C, key, accumulate, dtype, byte_off=0, name=""):
        lib = _lib.load()
        vec = 8 if dtype == BF16 else 4
        fold = 1
        while (C * fold) % vec:
            fold *= 2
        prow, pC = rows // fold, C * fold
        if pC // vec > 1024:
            # effdet_colsum gives every thread of a block one column vector (csrc/train.cu launch_colreduce): reject
            # here, when the plan is built, instead of at the first step.  Reached by the class head (C = 9 *
            # num_classes) unless its bias gradient comes out of the tensor-core weight-gradient launch (bf16, W <= 64):
            # num_classes <= 455 if a multiple of 4, <= 227 if even, <= 113 otherwise
            raise ValueError("bias gradient of %s: %d columns (x%d rows folded) exceed 1024 %d-wide column vectors "
                             "per block; for the class head use num_classes <= 113, an even count <= 227 or a "
                             "multiple of 4 <= 455" % (key, C, fold, vec))
        tail = rows - prow * fold            # < fold rows that do not fill a vector-aligned fold
        nblk = lib.effdet_colreduce_blocks(prow, pC, dtype)
        part = self._scratch(2 * pC * nblk, key + "_bg_partial")
        g = self.gw(key)
        self.add("bias_grad", [dz], [part],
                 lambda: _call("effdet_colsum", dz.ptr + byte_off, prow, pC, fold, g.data_ptr(), accumulate,
                               part.ptr, nblk, dtype), name)
        if tail:
            self.add("bias_grad", [dz], [part],
                     lambda: _call("effdet_colsum_tail", dz.ptr + byte_off, prow * fold, tail, C, g.data_ptr(),
                                   dtype), name + "_tail")

    # backward ------------------------------------------------------------------------
    def _build_backward(self):
        net, B, Wd = self.net, self.B, self.net.w_bifpn
        A = 9
        feats = self.pyramid
        hw = [f.shape[1] * f.shape[2] for f in feats]
        N = self.N
        lvl_off = np.concatenate([[0], np.cumsum([A * h for h in hw])[:-1]])
        # ---- heads
        for ht in self.head_tape:
            per = ht["per"]
            cout = A * per
            dz_final = self.dreg if ht["scope"] == "box_head" else self.dcls
            key = ht["final"]
            fuse_bias = self.use_tc_grads and self.wgrad_fuses_bias(Wd, 3)
            if not fuse_bias:
                # bias: the concatenated (B,N,per) tensor is a dense (B*N/9, 9*per) matrix
                self._bias_grad(dz_final, B * N // A, cout, key + "/bias", 0, F32, name=key + "_dbias")
            offs = [int(o) * per * 4 for o in lvl_off]
            layers = ht["layers"]
            targets = ht["xs"]
            gvals = [self.grad_of(t) for t in targets]
            if self.use_tc_grads:
                lv = self.lvl_dreg if ht["scope"] == "box_head" else self.lvl_dcls
                cpad = self.cpad_reg if ht["scope"] == "box_head" else self.cpad_cls
                self._wgrad_tc(ht["xs"], lv, key + "/kernel", Wd, cout, 3, dz_ld=cpad, name=key + "_wgrad",
                               bias_key=key + "/bias" if fuse_bias else None)
                wt = self._scratch(4, key + "_unused_T")
                # data gradient: K = padded channel count of the level buffers (padding is zero)
                self._dgrad_tc_padded(lv, key + "/kernel", cpad, cout, Wd, [g for g, _ in gvals], targets,
                                      masks=targets if layers else None, accumulate=[a for _, a in gvals],
                                      name=key + "_dgrad")
            else:
                self._wgrad(ht["xs"], [dz_final] * 5, key + "/kernel", Wd, cout, 3, 1,
                            dz_ld=[cout] * 5, dz_bs=[N * per] * 5, dz_off=offs, dz_dtype=F32,
                            name=key + "_wgrad")
                wt = self._transposed_weight(key + "/kernel", 9, Wd, cout)
                # data gradient into the last trunk layer's outputs (masked by its ReLU)
                self._dgrad(dz_final, None, wt, cout, Wd, [g for g, _ in gvals], targets,
                            masks=targets if layers else None, accumulate=[a for _, a in gvals],
                            x_ld=[cout] * 5, x_bs=[N * per] * 5, x_off=offs, in_dtype=F32,
                            shapes=[t.shape for t in targets], name=key + "_dgrad")
            for li in range(len(layers) - 1, -1, -1):
                L = layers[li]
                dzs = [self.gvals[id(y)] for y in L["ys"]]
                fuse_l = self.wgrad_fuses_bias(Wd, 3)       # bias gradient comes out of the wgrad launch
                if not fuse_l:
                    for l, dz in enumerate(dzs):
                        rows = B * dz.shape[1] * dz.shape[2]
                        self._bias_grad(dz, rows, Wd, L["name"] + "/bias", 0 if l == 0 else 1, self.dtype,
                                        name=L["name"] + "_dbias")
                self._wgrad(L["xs"], dzs, L["name"] + "/kernel", Wd, Wd, 3, 1, name=L["name"] + "_wgrad",
                            bias_key=L["name"] + "/bias" if fuse_l else None)
                wt = self._transposed_weight(L["name"] + "/kernel", 9, Wd, Wd)
                targets = L["xs"]
                gv = [self.grad_of(t) for t in targets]
                self._dgrad(None, dzs, wt, Wd, Wd, [g for g, _ in gv], targets,
                            masks=targets if li > 0 else None, accumulate=[a for _, a in gv],
                            shapes=[t.shape for t in targets], name=L["name"] + "_dgrad",
                            key=L["name"] + "/kernel")
        # gradient buckets (parallel.plan_buckets): (launches done, lowest flat index final) after each segment
        self.bucket_marks = [(len(self.ops), net.offsets["box_head/regress_head_conv_0/kernel"])]
        # ---- BiFPN (reverse tape)
        for rec in reversed(self.tape):
            if rec["kind"] == "node":
                self._node_backward(rec)
            else:
                self._convblock_backward(rec)
        self.bucket_marks.append((len(self.ops), net.backbone_end))
        # ---- backbone (only when it is trained): stages 7..5 hold ~85 % of its parameters and are
        # back-propagated first, so their bucket is reduced under the (long) backward of stages 4..1
        split = next((b.prefix for b in net.backbone.blocks if b.prefix.startswith("block5a")), None)
        for rec in reversed(self.bb_tape):
            if rec["kind"] == "mbconv":
                self._mbconv_backward(rec)
                if rec["blk"].prefix == split:
                    first = next(k for k in net.offsets if k.startswith(split))
                    self.bucket_marks.append((len(self.ops), net.offsets[first]))
            else:
                self._stem_backward(rec)
        if self.bb_tape:
            self.bucket_marks.append((len(self.ops), 0))
        self.bucket_marks[-1] = (len(self.ops), self.bucket_marks[-1][1])

    def _dgrad(self, x_single, xs, wt, cin, cout, dsts, targets, masks, accumulate, shapes, x_ld=None,
               x_bs=None, x_off=None, in_dtype=None, name="", key=None, extra_residual=None):
        """stride-1 3x3/1x1 data gradient = convolution of dz with the transposed kernel.
        cin/cout are those of THIS convolution (= forward Cout/Cin)."""
        n = len(dsts)
        in_dt = self.dtype if in_dtype is None else in_dtype
        taps = wt.shape[0] // (cin * cout)
        k = 3 if taps == 9 else 1
        use_tc = (self.net.use_tensor_cores and in_dt == BF16 and self.dtype == BF16 and key is not None
                  and cin % 8 == 0)
        panel = self.panel_for(key, taps, cout, cin, 1) if use_tc else None

        def make():
            d = _lib.ConvDesc()
            d.n_groups = n
            for i in range(n):
                src = x_single if x_single is not None else xs[i]
                d.x[i] = src.ptr + (x_off[i] if x_off else 0)
                d.y[i] = dsts[i].ptr
                d.residual[i] = dsts[i].ptr if accumulate[i] else (
                    extra_residual.ptr if extra_residual is not None else None)
                d.relu_mask[i] = masks[i].ptr if masks else None
                d.H[i], d.W[i] = shapes[i][1], shapes[i][2]
                d.ldx[i] = x_ld[i] if x_ld else 0
                d.x_batch_stride[i] = x_bs[i] if x_bs else 0
            d.B, d.Cin, d.Cout, d.kh, d.kw, d.stride = self.B, cin, cout, k, k, 1
            d.weight = wt.ptr
            d.act, d.in_dtype, d.out_dtype = ACT_NONE, in_dt, self.dtype
            if panel is not None:
                d.weight_bf16 = panel.ptr if isinstance(panel, Val) else panel.data_ptr()
                d.allow_tensor_core = 1
            else:
                d.allow_tensor_core = 0
            self._keepalive.append(d)
            return _call("effdet_conv2d", ctypes.byref(d))
        ins = ([x_single] if x_single is not None else list(xs)) + [wt] + (list(masks) if masks else [])
        ins += [dsts[i] for i in range(n) if accumulate[i]]
        if extra_residual is not None:
            ins.append(extra_residual)
        if isinstance(panel, Val):
            ins.append(panel)
        flops = sum(2 * self.B * s[1] * s[2] * cin * cout * taps for s in shapes)
        self.ops.append(Op("conv_dgrad_tc" if use_tc else "conv_dgrad", ins, list(dsts), make, name,
                           sum(v.nbytes for v in ins) + sum(v.nbytes for v in dsts), flops))

    def _dgrad_tc_padded(self, dzs, key, cpad, cout_fwd, cin_fwd, dsts, targets, masks, accumulate, name=""):
        """Data gradient of a head's final 3x3 conv from the channel-padded bf16 level buffers."""
        n = len(dsts)
        panel = self.panel_for(key, 9, cin_fwd, cout_fwd, 1)

        def make():
            d = _lib.ConvDesc()
            d.n_groups = n
            for i in range(n):
                d.x[i], d.y[i] = dzs[i].ptr, dsts[i].ptr
                d.residual[i] = dsts[i].ptr if accumulate[i] else None
                d.relu_mask[i] = masks[i].ptr if masks else None
                d.H[i], d.W[i] = targets[i].shape[1], targets[i].shape[2]
            # Cin = padded channel count: the panel's K extent (Cout_fwd rounded to 64) has zeros
            # beyond Cout_fwd and the buffers have zeros in their padding channels
            d.B, d.Cin, d.Cout, d.kh, d.kw, d.stride = self.B, cpad, cin_fwd, 3, 3, 1
            d.weight = None
            d.act, d.in_dtype, d.out_dtype = ACT_NONE, BF16, BF16
            d.weight_bf16 = panel.ptr if isinstance(panel, Val) else panel.data_ptr()
            d.allow_tensor_core = 1
            self._keepalive.append(d)
            return _call("effdet_conv2d", ctypes.byref(d))
        ins = list(dzs) + (list(masks) if masks else []) + [dsts[i] for i in range(n) if accumulate[i]]
        if isinstance(panel, Val):
            ins.append(panel)
        flops = sum(2 * self.B * t.shape[1] * t.shape[2] * cout_fwd * cin_fwd * 9 for t in targets)
        self.ops.append(Op("conv_dgrad_tc", ins, list(dsts), make, name,
                           sum(v.nbytes for v in ins) + sum(v.nbytes for v in dsts), flops))


    # ------------------------------------------------------------------ backbone in training mode
    def drop_keep(self, blk):
        i = self.drop_index.get(blk.prefix)
        return None if i is None else self.drop_scales[i]

    def _stem(self):
        if self.drop_blocks:
            nb = len(self.drop_blocks)
            self.add("drop_mask", [], [],
                     lambda: _call("effdet_drop_connect_scales", self.drop_rates.data_ptr(), nb, self.B,
                                   self.drop_seed, self.drop_step.data_ptr(), self.drop_scales.data_ptr()),
                     "drop_connect_scales")
        if not self.train_backbone:
            return super()._stem()
        net, B, S = self.net, self.B, self.net.image_size
        H = (S + 1) // 2
        c0 = net.backbone.stem_filters
        z = self.val((B, H, H, c0), name="stem_z", keep=True)
        y = self.val((B, H, H, c0), name="stem", keep=True)
        ones, zeros = net.const_ones(c0), net.const_zeros(c0)
        if self.u8_input:
            # a trainable stem needs the normalised image again for its weight gradient: materialise it once
            # on the device (the upload stays 3 B/pixel)
            raw, lut = self.images, net.normalization_lut()
            self.images_u8 = raw
            self.images = self.val((B, S, S, 3), F32, "images_f32", keep=True)
            self.add("stem", [raw], [self.images],
                     lambda: _call("effdet_normalize_u8", raw.ptr, lut.data_ptr(), self.images.ptr,
                                   B * S * S * 3), "normalize_image")
        # raw conv: identity scale/shift; swish is applied after the batch-norm pass, so the
        # stem kernel's built-in swish cannot be used -> run it through the generic conv (Cin = 3)
        if self.dtype == BF16 and c0 in (32, 40, 48, 56, 64):
            self.add("stem", [self.images], [z],
                     lambda: _call("effdet_stem_conv_act", self.images.ptr, self.w("stem_conv/kernel").data_ptr(),
                                   ones.data_ptr(), zeros.data_ptr(), z.ptr, B, S, S, c0, ACT_NONE), "stem_conv",
                     flops=2 * 27 * B * H * H * c0)
        else:
            self.conv([self.images], [z], "stem_conv/kernel", 3, c0, k=3, stride=2, in_dtype=F32, name="stem_conv")
        rec = self._bn_train(z, y, "stem_bn", c0, B * H * H, ACT_SWISH)
        rec.update(kind="stem", z=z, y=y, H=H, c0=c0)
        self.bb_tape.append(rec)
        return y, H

    def _mbconv(self, x, blk, H):
        if not self.train_backbone:
            return super()._mbconv(x, blk, H)
        B, net = self.B, self.net
        lib = _lib.load()
        p = blk.prefix
        inp, cin, cmid, cout = x, blk.input_filters, blk.mid_filters, blk.output_filters
        rec = dict(kind="mbconv", blk=blk, inp=inp, H=H)
        xin = x
        if blk.expand_ratio != 1:
            z_e = self.val((B, H, H, cmid), name=p + "expand_z", keep=True)
            y_e = self.val((B, H, H, cmid), name=p + "expand", keep=True)
            self.conv([x], [z_e], p + "expand_conv/kernel", cin, cmid, name=p + "expand_conv")
            rec["bn_e"] = self._bn_train(z_e, y_e, p + "expand_bn", cmid, B * H * H, ACT_SWISH)
            rec.update(z_e=z_e, y_e=y_e)
            xin = y_e
        Ho = (H + blk.stride - 1) // blk.stride
        z_d = self.val((B, Ho, Ho, cmid), name=p + "dw_z", keep=True)
        y_d = self.val((B, Ho, Ho, cmid), name=p + "dw", keep=True)
        # depthwise conv (raw) + batch statistics of its BN in one pass (bf16), then BN-apply + swish
        rec["bn_d"] = self._bn_train(z_d, y_d, p + "bn", cmid, B * Ho * Ho, ACT_SWISH,
                                     dw_from=(xin, p + "dwconv/depthwise_kernel", H, blk.kernel_size, blk.stride))
        HW = Ho * Ho
        nblk = lib.effdet_se_backward_blocks(HW, cmid, self.dtype)
        part = self.val((B, nblk, cmid), F32, name=p + "se_partial", keep=True)
        self.add("se_squeeze", [y_d], [part],
                 lambda: _call("effdet_spatial_sum", y_d.ptr, part.ptr, nblk, B, HW, cmid, self.dtype),
                 p + "se_squeeze")
        gate = self.val((B, cmid), F32, name=p + "gate", keep=True)
        self.add("se", [part], [gate],
                 lambda: _call("effdet_se_gate", part.ptr, nblk, 1.0 / float(HW),
                               self.w(p + "se_reduce/kernel").data_ptr(), self.w(p + "se_reduce/bias").data_ptr(),
                               self.w(p + "se_expand/kernel").data_ptr(), self.w(p + "se_expand/bias").data_ptr(),
                               gate.ptr, B, cmid, blk.se_filters), p + "se")
        yg = self.val((B, Ho, Ho, cmid), name=p + "se_excite", keep=True)
        self.add("se_apply", [y_d, gate], [yg],
                 lambda: _call("effdet_se_apply", y_d.ptr, gate.ptr, yg.ptr, B, HW, cmid, self.dtype),
                 p + "se_excite")
        z_p = self.val((B, Ho, Ho, cout), name=p + "project_z", keep=True)
        y_p = self.val((B, Ho, Ho, cout), name=p + "project", keep=True)
        self.conv([yg], [z_p], p + "project_conv/kernel", cmid, cout, name=p + "project_conv")
        rec["bn_p"] = self._bn_train(z_p, y_p, p + "project_bn", cout, B * HW, ACT_NONE)
        out = y_p
        if blk.has_skip:
            out = self.val((B, Ho, Ho, cout), name=p + "add", keep=True)
            ptrs = lambda: (ctypes.c_void_p * 3)(y_p.ptr, inp.ptr, None)

            def make_add():
                arr = ptrs()
                self._keepalive.append(arr)
                return _call("effdet_wbifpn_add", arr, 2, None, 0.0, out.ptr, B * HW * cout, self.dtype)
            keep = self.drop_keep(blk)
            if keep is not None:        # y_p * keep[b] / (1 - rate) + inp  (FixedDropout on the projected branch)
                self.add("add", [y_p, inp], [out],
                         lambda: _call("effdet_sample_scale_add", y_p.ptr, keep.data_ptr(), inp.ptr, out.ptr, B,
                                       HW * cout, self.dtype), p + "drop_add")
            else:
                self.add("add", [y_p, inp], [out], make_add, p + "add")
        rec.update(xin=xin, z_d=z_d, y_d=y_d, part=part, nblk=nblk, gate=gate, yg=yg, z_p=z_p, y_p=y_p,
                   out=out, Ho=Ho)
        self.bb_tape.append(rec)
        return out, Ho

    def _bn_act_backward(self, rec, dy, name):
        """dy (gradient of act(BN(z))) -> dz, for act in {none, relu, swish}."""
        z, C, rows, bn, act = rec["z"], rec["C"], rec["rows"], rec["bn"], rec["act"]
        dz = self.val(z.shape, name=name + "_dz")
        k123 = self._scratch(3 * C, bn + "/k123")
        w = self.w
        if rec["train"]:
            nblk = rec["nblk"]
            part = self._scratch(2 * C * nblk, bn + "/bwd_partial")
            mu, iv, ua, ub = rec["mean"], rec["invstd"], rec["ua"], rec["ub"]
            self.add("bn_bwd", [dy, z, mu, iv, ua, ub], [dz, k123, part],
                     lambda: _call("effdet_bn_act_backward", dy.ptr, z.ptr, rows, C, w(bn + "/gamma").data_ptr(),
                                   mu.ptr, iv.ptr, ua.ptr, ub.ptr, 0, act, self.gw(bn + "/gamma").data_ptr(),
                                   self.gw(bn + "/beta").data_ptr(), dz.ptr, k123.ptr, part.ptr, nblk,
                                   self.dtype), bn + "_bwd")
        else:
            fs, fb = self.folded(bn)
            self.add("bn_bwd", [dy, z], [dz, k123],
                     lambda: _call("effdet_bn_act_backward", dy.ptr, z.ptr, rows, C, w(bn + "/gamma").data_ptr(),
                                   None, None, fs.data_ptr(), fb.data_ptr(), 1, act, None, None, dz.ptr,
                                   k123.ptr, None, 1, self.dtype), bn + "_bwd")
        return dz

    def _mbconv_backward(self, rec):
        lib = _lib.load()
        B, blk = self.B, rec["blk"]
        p = blk.prefix
        cin, cmid, cout, H, Ho = blk.input_filters, blk.mid_filters, blk.output_filters, rec["H"], rec["Ho"]
        HW = Ho * Ho
        d_out = self.gvals.get(id(rec["out"]))
        if d_out is None:
            return
        # project BN (linear) -> project conv
        bn_p = dict(rec["bn_p"]); bn_p["z"] = rec["z_p"]
        d_yp = d_out
        keep = self.drop_keep(blk) if blk.has_skip else None
        if keep is not None:            # backward of the dropped branch: d y_p = d out * keep[b] / (1 - rate)
            d_yp = self.val(rec["y_p"].shape, name=p + "drop_grad")
            self.add("add", [d_out], [d_yp],
                     lambda: _call("effdet_sample_scale_add", d_out.ptr, keep.data_ptr(), None, d_yp.ptr, B,
                                   HW * cout, self.dtype), p + "drop_bwd")
        dz_p = self._bn_act_backward(bn_p, d_yp, p + "project")
        pkey = p + "project_conv/kernel"
        yg = rec["yg"]
        self._wgrad([yg], [dz_p], pkey, cmid, cout, 1, 1, name=p + "project_wgrad")
        dyg = self.val(yg.shape, name=p + "se_excite_grad")
        wt = self._transposed_weight(pkey, 1, cmid, cout)
        self._dgrad(None, [dz_p], wt, cout, cmid, [dyg], [yg], None, [0], [yg.shape], name=p + "project_dgrad",
                    key=pkey)
        # squeeze-excite, then depthwise BN + swish
        y_d, gate, part, nblk = rec["y_d"], rec["gate"], rec["part"], rec["nblk"]
        R = blk.se_filters
        fcs = self._scratch(B * (2 * cmid * R + R + cmid), p + "se_fc_scratch")
        dmean = self._scratch(B * cmid, p + "se_dmean")
        w, gw = self.w, self.gw
        bn_d = dict(rec["bn_d"]); bn_d["z"] = rec["z_d"]
        if bn_d["train"] and bn_d["act"] == ACT_SWISH and os.environ.get("EFFDET_SE_BN_FUSED", "1") != "0":
            # SE gate backward + depthwise BN / swish backward in one reduction and one apply pass over
            # (dyg, z_d): y_d and swish' are recomputed from z_d, dy_d is never written (effdet_se_bn_backward:
            # five instead of nine passes over the expanded tensors)
            z_d, bn = rec["z_d"], bn_d["bn"]
            dz_d = self.val(z_d.shape, name=p + "dw_dz")
            nb2 = lib.effdet_se_bn_backward_blocks(B, HW, cmid, self.dtype)
            dgp2 = self._scratch(B * (nb2 + 1) * cmid, p + "se_dg_partial")
            bnp = self._scratch(B * (nb2 + 1) * 4 * cmid, p + "se_bn_partial")
            bnr = self._scratch(B * 2 * cmid, p + "se_bn_rows")
            k123 = self._scratch(3 * cmid, bn + "/k123")
            mu, iv, ua, ub = bn_d["mean"], bn_d["invstd"], bn_d["ua"], bn_d["ub"]
            self.add("se_bn_bwd", [dyg, z_d, gate, part, mu, iv, ua, ub], [dz_d, k123, dgp2, bnp, bnr, fcs, dmean],
                     lambda: _call("effdet_se_bn_backward", dyg.ptr, z_d.ptr, gate.ptr, part.ptr, nblk,
                                   w(p + "se_reduce/kernel").data_ptr(), w(p + "se_reduce/bias").data_ptr(),
                                   w(p + "se_expand/kernel").data_ptr(), w(p + "se_expand/bias").data_ptr(),
                                   gw(p + "se_reduce/kernel").data_ptr(), gw(p + "se_reduce/bias").data_ptr(),
                                   gw(p + "se_expand/kernel").data_ptr(), gw(p + "se_expand/bias").data_ptr(),
                                   w(bn + "/gamma").data_ptr(), mu.ptr, iv.ptr, ua.ptr, ub.ptr,
                                   gw(bn + "/gamma").data_ptr(), gw(bn + "/beta").data_ptr(), dz_d.ptr, k123.ptr,
                                   dgp2.ptr, bnp.ptr, bnr.ptr, nb2, fcs.ptr, dmean.ptr, B, HW, cmid, R, self.dtype),
                     p + "se_bn_bwd")
        else:
            dy_d = self.val(y_d.shape, name=p + "dw_grad")
            dgb = lib.effdet_se_backward_blocks(HW, cmid, self.dtype)
            dgp = self._scratch(B * dgb * cmid, p + "se_dg_partial")
            self.add("se_bwd", [dyg, y_d, gate, part], [dy_d, dgp, fcs, dmean],
                     lambda: _call("effdet_se_backward", dyg.ptr, y_d.ptr, gate.ptr, part.ptr, nblk,
                                   w(p + "se_reduce/kernel").data_ptr(), w(p + "se_reduce/bias").data_ptr(),
                                   w(p + "se_expand/kernel").data_ptr(), w(p + "se_expand/bias").data_ptr(),
                                   dy_d.ptr, gw(p + "se_reduce/kernel").data_ptr(),
                                   gw(p + "se_reduce/bias").data_ptr(), gw(p + "se_expand/kernel").data_ptr(),
                                   gw(p + "se_expand/bias").data_ptr(), dgp.ptr, dgb, fcs.ptr, dmean.ptr, B, HW,
                                   cmid, R, self.dtype), p + "se_bwd")
            dz_d = self._bn_act_backward(bn_d, dy_d, p + "dw")
        xin = rec["xin"]
        k, st = blk.kernel_size, blk.stride
        nb = lib.effdet_dw_backward_blocks(B, H, H, cmid, k, st, self.dtype)
        dwp = self._scratch(k * k * cmid * nb, p + "dw_partial")
        dkey = p + "dwconv/depthwise_kernel"
        inp = rec["inp"]
        if blk.expand_ratio != 1:
            d_xin = self.val(xin.shape, name=p + "expand_grad")
        else:
            d_xin, acc = self.grad_of(inp)
            assert acc == 0, "block without expansion must be the only consumer of its input"
        if st == 2 and self.dtype == BF16:
            # stride-2 data gradient on the TMA tile pipeline: zero-insert dz into an x-sized tensor, then
            # the stride-1 depthwise kernel with spatially flipped taps (the gather-form kernel that
            # effdet_dw_backward falls back to runs at 0.5-1.4 TB/s); the weight gradient stays there
            self.add("dw_bwd", [xin, dz_d], [dwp],
                     lambda: _call("effdet_dw_backward", xin.ptr, dz_d.ptr, w(dkey).data_ptr(), None,
                                   gw(dkey).data_ptr(), dwp.ptr, nb, B, H, H, cmid, k, st, self.dtype),
                     p + "dw_wgrad", flops=2 * k * k * B * Ho * Ho * cmid)
            self._dw_dgrad_stride2(dz_d, dkey, d_xin, H, Ho, cmid, k, p)
        else:
            self.add("dw_bwd", [xin, dz_d], [d_xin, dwp],
                     lambda: _call("effdet_dw_backward", xin.ptr, dz_d.ptr, w(dkey).data_ptr(), d_xin.ptr,
                                   gw(dkey).data_ptr(), dwp.ptr, nb, B, H, H, cmid, k, st, self.dtype),
                     p + "dw_bwd", flops=4 * k * k * B * Ho * Ho * cmid)
        if blk.expand_ratio == 1:
            if blk.has_skip:
                # no expand conv whose data-gradient epilogue could carry the residual branch (block1b.. of
                # B1-B6): d(inp) = depthwise data gradient + d(out)   (efficientnet.py:297-304)
                def make_skip_add():
                    arr = (ctypes.c_void_p * 3)(d_xin.ptr, d_out.ptr, None)
                    self._keepalive.append(arr)
                    return _call("effdet_wbifpn_add", arr, 2, None, 0.0, d_xin.ptr, B * H * H * cin, self.dtype)
                self.add("add", [d_xin, d_out], [d_xin], make_skip_add, p + "skip_grad_add")
            return
        bn_e = dict(rec["bn_e"]); bn_e["z"] = rec["z_e"]
        dz_e = self._bn_act_backward(bn_e, d_xin, p + "expand")
        ekey = p + "expand_conv/kernel"
        self._wgrad([inp], [dz_e], ekey, cin, cmid, 1, 1, name=p + "expand_wgrad")
        if not self._needs_grad(inp):
            return
        g, acc = self.grad_of(inp)
        extra = None
        if blk.has_skip:
            if acc:
                raise NotImplementedError("skip block whose input already has a gradient")
            extra = d_out                     # residual branch: d(inp) = d(out) + dgrad(expand)
        wt = self._transposed_weight(ekey, 1, cin, cmid)
        self._dgrad(None, [dz_e], wt, cmid, cin, [g], [inp], None, [acc], [inp.shape], name=p + "expand_dgrad",
                    key=ekey, extra_residual=extra)

    def _dw_dgrad_stride2(self, dz, dkey, dx, H, Ho, C, k, p):
        """dx (B,H,H,C) of a stride-2 SAME depthwise conv (efficientnet.py:242-252).  With the forward
        iy = 2*oy - pad_t + ky, dx[iy] = sum_ky U[iy + pad_t - ky] w[ky] for U = dz with zeros between samples;
        the stride-1 SAME conv with flipped taps computes sum_ky V[iy + (k-1)/2 - ky] w[ky], so V is U
        shifted by (k-1)/2 - pad_t (1 for even H, 0 for odd H)."""
        B = self.B
        pad_t = max((Ho - 1) * 2 + k - H, 0) // 2
        shift = (k - 1) // 2 - pad_t
        assert shift in (0, 1) and 2 * (Ho - 1) + shift < H
        U = self.val((B, H, H, C), name=p + "dw_dz_zero_inserted")
        self.add("zero_insert", [dz], [U],
                 lambda: _call("effdet_zero_insert", dz.ptr, U.ptr, B, Ho, Ho, C, H, H, shift, shift, self.dtype),
                 p + "dw_zero_insert")
        wflip = self._scratch(k * k * C, p + "dw_flip")
        self.add("wtrans", [], [wflip],
                 lambda: _call("effdet_flip_taps", self.w(dkey).data_ptr(), wflip.ptr, k * k, C), p + "dw_flip")
        ones, zeros = self.net.const_ones(C), self.net.const_zeros(C)
        self.add("dwconv", [U, wflip], [dx],
                 lambda: _call("effdet_dwconv", U.ptr, wflip.ptr, ones.data_ptr(), zeros.data_ptr(), dx.ptr,
                               None, 0, B, H, H, C, k, 1, ACT_NONE, self.dtype), p + "dw_dgrad_s2",
                 flops=2 * k * k * B * H * H * C)

    def _stem_backward(self, rec):
        lib = _lib.load()
        B, S = self.B, self.net.image_size
        dy = self.gvals.get(id(rec["y"]))
        if dy is None:
            return
        dz = self._bn_act_backward(rec, dy, "stem")
        nb = lib.effdet_stem_wgrad_blocks(B, S, S)
        c0 = rec["c0"]
        part = self._scratch(27 * c0 * nb, "stem_wg_partial")
        self.add("stem_wgrad", [self.images, dz], [part],
                 lambda: _call("effdet_stem_wgrad", self.images.ptr, dz.ptr, self.gw("stem_conv/kernel").data_ptr(),
                               part.ptr, nb, B, S, S, c0, self.dtype), "stem_wgrad")

    def _bn_backward(self, rec, dy):
        """dy (grad of y = relu(BN(z))) -> dz (new Val)."""
        z, y, C, rows, bn = rec["z"], rec["y"], rec["C"], rec["rows"], rec["bn"]
        dz = self.val(z.shape, name=(z.name or "") + "_grad")
        k123 = self._scratch(3 * C, bn + "/k123")
        w = self.w
        if rec["train"]:
            nblk = rec["nblk"]
            part = self._scratch(2 * C * nblk, bn + "/bwd_partial")
            mu, iv = rec["mean"], rec["invstd"]
            self.add("bn_bwd", [dy, y, z, mu, iv], [dz, k123, part],
                     lambda: _call("effdet_bn_relu_backward", dy.ptr, y.ptr, z.ptr, rows, C,
                                   w(bn + "/gamma").data_ptr(), mu.ptr, iv.ptr, None,
                                   self.gw(bn + "/gamma").data_ptr(), self.gw(bn + "/beta").data_ptr(),
                                   dz.ptr, k123.ptr, part.ptr, nblk, self.dtype), bn + "_bwd")
        else:
            fs, _ = self.folded(bn)
            part = self._scratch(4, bn + "/bwd_partial")
            self.add("bn_bwd", [dy, y, z], [dz, k123],
                     lambda: _call("effdet_bn_relu_backward", dy.ptr, y.ptr, z.ptr, rows, C,
                                   w(bn + "/gamma").data_ptr(), None, None, fs.data_ptr(), None, None,
                                   dz.ptr, k123.ptr, part.ptr, 1, self.dtype), bn + "_bwd")
        return dz

    def _needs_grad(self, v):
        if self.train_backbone:
            return v is not self.images
        return not any(v is f for f in self.features)     # frozen backbone features

    def _node_backward(self, rec):
        lib = _lib.load()
        B, H, C, name = self.B, rec["H"], rec["C"], rec["name"]
        net = self.net
        y = rec["y"]
        dy = self.gvals.get(id(y))
        if dy is None:
            return
        dz = self._bn_backward(rec, dy)
        f = rec["f"]
        kkey = name + "_dconv/depthwise_kernel"
        nblk = lib.effdet_dw_wgrad_blocks(B, H, H, C, self.dtype)
        part = self._scratch(9 * C * nblk, name + "/dw_wg_partial")
        self.add("dw_wgrad", [f, dz], [part],
                 lambda: _call("effdet_dw_wgrad", f.ptr, dz.ptr, B, H, H, C, self.gw(kkey).data_ptr(),
                               part.ptr, nblk, self.dtype), name + "_dw_wgrad", flops=18 * B * H * H * C)
        wflip = self._scratch(9 * C, name + "/dw_flip")
        self.add("wtrans", [], [wflip],
                 lambda: _call("effdet_flip_taps", self.w(kkey).data_ptr(), wflip.ptr, 9, C), name + "_flip")
        df = self.val(f.shape, name=name + "_f_grad")
        ones, zeros = net.const_ones(C), net.const_zeros(C)
        self.add("dwconv", [dz, wflip], [df],
                 lambda: _call("effdet_dwconv", dz.ptr, wflip.ptr, ones.data_ptr(), zeros.data_ptr(), df.ptr,
                               None, 0, B, H, H, C, 3, 1, ACT_NONE, self.dtype), name + "_dw_dgrad",
                 flops=18 * B * H * H * C)
        in0, in1, in2, mode0 = rec["in0"], rec["in1"], rec["in2"], rec["mode0"]
        n_in = 3 if in2 is not None else 2
        fw = self.w(rec["fuse"] + "/" + rec["fuse"]) if net.weighted_bifpn else None
        fwp = fw.data_ptr() if fw is not None else None
        if fw is not None:
            fpart = self._scratch(4 * 148 * 8, name + "/fuse_wg_partial")
            gfw = self.gw(rec["fuse"] + "/" + rec["fuse"])
            self.add("fuse_wgrad", [df, f, in0, in1, in2], [fpart],
                     lambda: _call("effdet_fuse_backward_weights", df.ptr, f.ptr, in0.ptr, mode0, in1.ptr,
                                   in2.ptr if in2 is not None else None, fwp, 1e-4, gfw.data_ptr(),
                                   fpart.ptr, B, H, H, C, self.dtype), name + "_fuse_wgrad")
        for which, src in enumerate([in0, in1, in2][:n_in]):
            if not self._needs_grad(src):
                continue
            g, acc = self.grad_of(src)
            self.add("fuse_bwd", [df, in0 if (which == 0 and mode0 == DOWN) else None, g if acc else None],
                     [g],
                     lambda which=which, g=g, acc=acc:
                     _call("effdet_fuse_backward_input", df.ptr, which, mode0, in0.ptr, fwp, n_in, 1e-4,
                           g.ptr, acc, B, H, H, C, self.dtype), name + "_fuse_bwd%d" % which)

    def _convblock_backward(self, rec):
        B = self.B
        y, x = rec["y"], rec["x"]
        dy = self.gvals.get(id(y))
        if dy is None:
            return
        dz = self._bn_backward(rec, dy)
        name, cin, cout, k, stride = rec["name"], rec["cin"], rec["cout"], rec["k"], rec["stride"]
        key = name + "_conv/kernel"
        self._wgrad([x], [dz], key, cin, cout, k, stride, name=name + "_wgrad")
        if not self._needs_grad(x):
            return
        g, acc = self.grad_of(x)
        if stride == 1:
            wt = self._transposed_weight(key, k * k, cin, cout)
            self._dgrad(None, [dz], wt, cout, cin, [g], [x], None, [acc], [x.shape], name=name + "_dgrad",
                        key=key)
        elif self.net.use_tensor_cores and self.dtype == BF16 and stride == 2 and cout % 8 == 0:
            # stride 2: zero-insert dz to the input's extent, then the stride-1 tensor-core data gradient
            H, Ho = x.shape[1], dz.shape[1]
            a = (k - 1) // 2 - max((Ho - 1) * 2 + k - H, 0) // 2
            dzs = self.val((B, H, H, cout), name=name + "_dz_up")
            self.add("zero_insert", [dz], [dzs],
                     lambda: _call("effdet_zero_insert", dz.ptr, dzs.ptr, B, Ho, Ho, cout, H, H, a, a, self.dtype),
                     name + "_dz_up")
            wt = self._transposed_weight(key, k * k, cin, cout)
            self._dgrad(None, [dzs], wt, cout, cin, [g], [x], None, [acc], [x.shape], name=name + "_dgrad",
                        key=key)
        else:
            H = x.shape[1]
            self.add("conv_dgrad", [dz, g if acc else None], [g],
                     lambda: _call("effdet_conv_dgrad_strided", dz.ptr, self.w(key).data_ptr(), g.ptr, acc, B,
                                   H, H, cin, cout, k, stride, self.dtype), name + "_dgrad")


class Trainer:
    """compile()/train_on_batch() backend of model.Model."""

    def __init__(self, model, optimizer, loss):
        from .optimizers import SGD
        from .utils.tpu import Focal, SmoothL1
        self.model, self.net = model, model.net
        self.opt = optimizer if optimizer is not None else SGD(lr=0.01, decay=4e-5, momentum=0.9)
        loss = loss or {}
        self.focal = loss.get("classification") or Focal(0.25, 1.5)
        self.sl1 = loss.get("regression") or SmoothL1(1)
        self.plans = {}
        self.world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        self.graphs = {}

    def plan(self, B, dense, u8=False):
        key = (B, dense, bool(u8))
        if key not in self.plans:
            net = self.net
            names = net.backbone.keras_layer_names()
            frozen = net.frozen_layers
            n_frozen = sum(1 for n in names if n in frozen)
            if 0 < n_frozen < len(names):
                raise NotImplementedError(
                    "partially frozen backbones are not supported: freeze all of it "
                    "(model.freeze_backbone(), train_tpu.py --freeze-backbone) or none of it")
            self.train_backbone = n_frozen == 0
            self.plans[key] = TrainPlan(net, B, self.focal.alpha, self.focal.gamma, self.sl1.lambda_,
                                        dense_labels=dense,
                                        reuse_buffers=os.environ.get("EFFDET_NO_REUSE") != "1",
                                        train_backbone=self.train_backbone, u8_input=bool(u8))
        return self.plans[key]

    def load_batch(self, plan, images, targets):
        """targets: [regression_t (B,N,5), labels_t (B,N,C+1)] (reference generator layout) or
        (regression_t, state i8, cls i32) device tensors from anchor_targets_device(compact=True)."""
        dev = self.net.device
        img = images if isinstance(images, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(images, np.uint8 if plan.u8_input else np.float32))
        plan.tensor(plan.input_images).copy_(img.to(dev, non_blocking=True))
        if plan.dense_labels:
            reg_t, lab_t = targets
            plan.tensor(plan.reg_t).copy_(torch.as_tensor(reg_t).to(dev, non_blocking=True))
            plan.tensor(plan.lab_t).copy_(torch.as_tensor(lab_t).to(dev, non_blocking=True))
        else:
            reg_t, st, cl = targets
            plan.tensor(plan.reg_t).copy_(reg_t)
            plan.tensor(plan.state_t).copy_(st)
            plan.tensor(plan.cls_t).copy_(cl)

    def _reduce_and_update(self, lo, hi):
        """flat[lo:hi): zero the gradients of frozen layers, SUM over replicas (NCCL over NVLink; the 1/replicas
        factor rides in the SGD kernel), SGD-momentum update -- all on the current stream."""
        net = self.net
        for k in self._frozen_keys():          # layers.trainable = False outside the backbone
            if lo <= net.offsets[k] < hi:
                net.grads[k].zero_()
        from . import parallel
        g = net.grad_flat[lo:hi]
        scale = parallel.allreduce_gradients_(g)
        _lib.call("effdet_sgd_momentum_step", net.flat.data_ptr() + 4 * lo, g.data_ptr(),
                  net.velocity.data_ptr() + 4 * lo, hi - lo, float(self.opt.current_lr()),
                  float(self.opt.momentum), float(scale), _lib.stream_ptr(net.device))

    def apply_gradients(self):
        net = self.net
        start = 0 if getattr(self, "train_backbone", False) else net.backbone_end
        self._reduce_and_update(start, net.flat.numel())
        self.opt.iterations += 1
        net.invalidate()            # folded BN / static weight panels of the inference plans are stale now

    def run_step(self, plan):
        """One optimizer step on the batch already loaded into `plan`: forward + losses + backward (captured CUDA
        graph) -> gradient all-reduce -> SGD.  With several replicas the backward is captured as segments
        (heads | BiFPN | backbone stages 7-5 | stages 4-1 + stem) and each segment's gradient bucket is
        all-reduced and applied on a second stream while the next segment computes -- the bucketed, overlapped
        form of MirroredStrategy's per-step all-reduce (train_tpu.py:233-247)."""
        if plan.graph is None:
            plan.capture()
        buckets = getattr(plan, "buckets", None)
        if not buckets or len(buckets) == 1:
            plan.replay()
            self.apply_gradients()
            return
        net, dev = self.net, self.net.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_comm_stream", None) is None:
            self._comm_stream = torch.cuda.Stream(dev)
            self._seg_events = [torch.cuda.Event() for _ in range(8)]
        comm = self._comm_stream
        start = 0 if getattr(self, "train_backbone", False) else net.backbone_end
        for i, (_, lo, hi) in enumerate(buckets):
            plan.replay_segment(i)
            lo = max(lo, start)
            if hi <= lo:
                continue
            ev = self._seg_events[i]
            ev.record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                self._reduce_and_update(lo, hi)
        main.wait_stream(comm)          # the next forward reads the updated weights
        self.opt.iterations += 1
        net.invalidate()

    def _frozen_keys(self):
        net = self.net
        fl = net.frozen_layers
        keys = []
        for k in net.offsets:
            if net.offsets[k] < net.backbone_end or k.startswith("boxes/"):
                continue
            layer = k.rsplit("/", 1)[0]
            top = layer.split("/")[0]
            if layer in fl or top in fl or (k.startswith("w_bi_fpn_add") and top in fl):
                keys.append(k)
        return keys

    def targets_into_plan(self, plan, anchors_d, gt_d, gl_d, cnt_d, hw_d, kmax):
        """Row 12 on the device: writes regression / compact class targets straight into the
        plan's buffers (no dense one-hot tensor, no host round trip)."""
        net = self.net
        _lib.call("effdet_anchor_targets", anchors_d.data_ptr(), plan.N, gt_d.data_ptr(), gl_d.data_ptr(),
                  cnt_d.data_ptr(), plan.B, int(kmax), hw_d.data_ptr(), net.num_classes, 0.4, 0.5,
                  plan.reg_t.ptr, None, plan.state_t.ptr, plan.cls_t.ptr, _lib.stream_ptr(net.device))

    def fit_prefetched(self, plan, anchors_d, batches):
        """Training loop over an iterable of HOST batches
        (images (B,S,S,3) f32, (boxes, labels, counts, image_hw) packed annotations, kmax), all pinned:
        the host->device copies of batch i+1 run on a copy stream while step i (device target
        assignment -> captured forward/backward graph -> all-reduce -> SGD) executes -- the
        dataset.prefetch() of the reference's input pipeline (train_tpu.py:186-228).
        Yields the (8,) loss record of every step as a host tensor."""
        dev = self.net.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
            self._slots = {}
        cs = self._copy_stream
        img_buf = plan.tensor(plan.input_images)

        def stage(slot, batch):
            imgs, gt, kmax = batch
            host = [imgs] + list(gt)
            key = (slot,) + tuple((tuple(t.shape), t.dtype) for t in host)
            if key not in self._slots:
                self._slots[key] = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host]
            d = self._slots[key]
            with torch.cuda.stream(cs):
                for dst, src in zip(d, host):
                    dst.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            return d, kmax, ev
        # The loss record of step i is read back while step i+1 is already enqueued: an asynchronous copy into a
        # pinned buffer right behind step i on the main stream, waited for one iteration later.  (Reading it
        # synchronously left the GPU idle between steps for the ~0.15 ms the host needs to enqueue the next one:
        # 4-5 % of a 3.8 ms D0 step.)  A staging slot is rewritten only after the step that consumed it has copied
        # its images / annotations out (`consumed` events), no longer implied by the per-step synchronisation.
        if getattr(self, "_loss_host", None) is None:
            self._loss_host = [torch.empty(plan.tensor(plan.loss_out).shape, dtype=torch.float32).pin_memory()
                               for _ in range(2)]
        consumed = [None, None]
        it = iter(batches)
        first = next(it, None)
        nxt = stage(0, first) if first is not None else None
        i = 0
        pending = None
        while nxt is not None:
            d, kmax, ev = nxt
            b = next(it, None)
            if b is not None:
                slot = (i + 1) % 2
                if consumed[slot] is not None:
                    cs.wait_event(consumed[slot])
                nxt = stage(slot, b)
            else:
                nxt = None
            main.wait_event(ev)
            img_buf.copy_(d[0], non_blocking=True)
            self.targets_into_plan(plan, anchors_d, d[1], d[2], d[3], d[4], kmax)
            cev = torch.cuda.Event()
            cev.record(main)
            consumed[i % 2] = cev
            self.run_step(plan)
            hb = self._loss_host[i % 2]
            hb.copy_(plan.tensor(plan.loss_out), non_blocking=True)
            lev = torch.cuda.Event()
            lev.record(main)
            if pending is not None:
                pending[1].synchronize()
                yield pending[0].clone()
            pending = (hb, lev)
            i += 1
        if pending is not None:
            pending[1].synchronize()
            yield pending[0].clone()

    def step(self, images, targets, sync=True):
        dense = not (isinstance(targets, (tuple, list)) and len(targets) == 3)
        B = int(images.shape[0])
        from .model import _is_u8
        plan = self.plan(B, dense, u8=_is_u8(images))     # uint8 = raw letterboxed RGB (utils.preprocess)
        self.load_batch(plan, images, targets)
        if os.environ.get("EFFDET_EAGER_STEP") == "1":
            plan.run()
            self.apply_gradients()
        else:
            self.run_step(plan)     # the captured graph(s): the path bench.py times (fit_prefetched)
        if not sync:
            return None
        out = plan.tensor(plan.loss_out).cpu().numpy()
        return [float(out[0] + out[1]), float(out[1]), float(out[0])]   # keras: [total, regression, classification]
