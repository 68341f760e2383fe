"""Oracle: the reference Keras graph restated on torch-CPU (TEST INFRASTRUCTURE --
see oracle/__init__.py).  WIRING PINNED: tests/test_oracle_graph_golden.py checks
this file against outputs of the reference's own model.py / efficientnet.py /
layers.py executed unmodified under a torch-backed Keras stand-in (D0, D0 weighted,
D1, D3 weighted: C3-C5, every BiFPN level, regression, classification to 2e-6).
The arithmetic inside TensorFlow's kernels is restated from SURVEY.md Appendix A
(each TF behaviour sits behind one function); TensorFlow itself cannot run here.

Follows
  /root/reference/efficientnet.py:99-114 (block table) :191-207 (rounding)
      :210-306 (mb_conv_block) :309-470 (EfficientNet -> 5 features) :473-575 (B0..B6)
  /root/reference/model.py:28-38 (scaling tables) :42-45 (BN config) :48-90 (blocks)
      :93-196 (BiFPN) :199-268 (wBiFPN) :271-353 (heads) :356-407 (assembly)
  /root/reference/layers.py:26-31 (fast normalised fusion)
Weights: dict keyed "<keras layer name>/<weight name>" with Keras shapes
(Conv kernel HWIO, depthwise kernel HWC1, BN gamma/beta/moving_mean/moving_variance);
activations are NHWC at the interface, NCHW inside (torch convention).
"""
import math

import numpy as np
import torch
import torch.nn.functional as Fn

W_BIFPNS = [64, 88, 112, 160, 224, 288, 384]              # model.py:28
IMAGE_SIZES = [512, 640, 768, 896, 1024, 1280, 1408]      # model.py:29
COEFFS = [(1.0, 1.0), (1.0, 1.1), (1.1, 1.2), (1.2, 1.4), (1.4, 1.8), (1.6, 2.2), (1.8, 2.6)]
# (kernel, repeats, in, out, expand, stride)  efficientnet.py:99-114 ; se_ratio .25, id_skip True
BLOCKS = [(3, 1, 32, 16, 1, 1), (3, 2, 16, 24, 6, 2), (5, 2, 24, 40, 6, 2), (3, 3, 40, 80, 6, 2),
          (5, 3, 80, 112, 6, 1), (5, 4, 112, 192, 6, 2), (3, 1, 192, 320, 6, 1)]
BN_EPS_BACKBONE = 1e-3        # keras BatchNormalization default (Appendix A.2)
BN_EPS_BIFPN = 1e-4           # model.py:42-45
BN_MOM_BIFPN = 0.997


def round_filters(filters, width, divisor=8):          # efficientnet.py:191-201
    filters *= width
    new = int(filters + divisor / 2) // divisor * divisor
    new = max(divisor, new)
    if new < 0.9 * filters:
        new += divisor
    return int(new)


def round_repeats(r, depth):                           # efficientnet.py:204-207
    return int(math.ceil(depth * r))


def block_list(phi):
    """[(prefix, k, stride, cin, cout, expand, se_dim, has_skip, drop_index)] + feature taps."""
    wc, dc = COEFFS[phi]
    out, taps = [], []
    num = 0
    for idx, (k, rep, cin, cout, e, s) in enumerate(BLOCKS):
        cin, cout, rep = round_filters(cin, wc), round_filters(cout, wc), round_repeats(rep, dc)
        for r in range(rep):
            ci = cin if r == 0 else cout
            st = s if r == 0 else 1
            out.append(dict(prefix="block%d%s_" % (idx + 1, "abcdefghijklmnopqrstuvwxyz"[r]),
                            k=k, stride=st, cin=ci, cout=cout, expand=e,
                            se=max(1, int(ci * 0.25)), skip=(st == 1 and ci == cout), num=num))
            num += 1
        if (idx < len(BLOCKS) - 1 and BLOCKS[idx + 1][5] == 2) or idx == len(BLOCKS) - 1:
            taps.append(len(out) - 1)
    return out, taps


# ---------------------------------------------------------------- TF op semantics
def same_pad(x, k, s):                                  # Appendix A.1
    H, W = x.shape[-2:]
    def p(n):
        o = -(-n // s)
        t = max((o - 1) * s + k - n, 0)
        return t // 2, t - t // 2
    (t, b), (l, r) = p(H), p(W)
    return Fn.pad(x, (l, r, t, b))


def conv2d(x, w, stride=1, bias=None):
    k = w.shape[0]
    wt = torch.as_tensor(w).permute(3, 2, 0, 1).to(x.dtype)
    b = None if bias is None else torch.as_tensor(bias).to(x.dtype)
    return Fn.conv2d(same_pad(x, k, stride), wt, b, stride=stride)


def dwconv2d(x, w, stride=1):
    k, C = w.shape[0], w.shape[2]
    wt = torch.as_tensor(w).permute(2, 3, 0, 1).to(x.dtype)          # (C,1,k,k)
    return Fn.conv2d(same_pad(x, k, stride), wt, None, stride=stride, groups=C)


def batchnorm(x, W, name, eps, training=False, stats_out=None):
    g = torch.as_tensor(W[name + "/gamma"]).to(x.dtype)
    b = torch.as_tensor(W[name + "/beta"]).to(x.dtype)
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if stats_out is not None:
            n = x.numel() // x.shape[1]
            stats_out[name] = (mean.detach(), var.detach() * (n / max(n - 1, 1)))
    else:
        mean = torch.as_tensor(W[name + "/moving_mean"]).to(x.dtype)
        var = torch.as_tensor(W[name + "/moving_variance"]).to(x.dtype)
    sh = (1, -1, 1, 1)
    return (x - mean.view(sh)) / torch.sqrt(var.view(sh) + eps) * g.view(sh) + b.view(sh)


def swish(x):
    return x * torch.sigmoid(x)


def forced(x, force, name):
    """Teacher forcing for the bf16 training-parity test: when `force` holds a tensor for `name` (NHWC, the value
    the CUDA path stored for this activation), the forward value becomes exactly that tensor while the gradient
    still flows to x (straight-through).  ReLU masks and max-pool routes of the backward pass are then those of
    the CUDA forward, so the remaining gradient difference is the backward arithmetic alone (a 2^-9 relative
    perturbation of a ReLU network's forward otherwise flips ~0.4 % of the masks per layer, which moves the
    gradients by tens of per cent in L2 -- tests/test_gpu_train.py)."""
    if force is None or name not in force:
        return x
    f = torch.as_tensor(force[name]).to(x.dtype).permute(0, 3, 1, 2)
    assert f.shape == x.shape, (name, f.shape, x.shape)
    return x + (f - x).detach()


def relu_forced(pre, force, name):
    """relu(pre); under teacher forcing the ReLU mask is the forced activation's (f > 0)."""
    if force is None or name not in force:
        return torch.relu(pre)
    f = torch.as_tensor(force[name]).to(pre.dtype).permute(0, 3, 1, 2)
    assert f.shape == pre.shape, (name, f.shape, pre.shape)
    y = pre * (f > 0).to(pre.dtype)
    return y + (f - y).detach()


def upsample2(x):                                       # UpSampling2D(): nearest x2
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def maxpool2(x):                                        # MaxPooling2D(strides=2): 2x2 valid
    return Fn.max_pool2d(x, 2, 2)


# ---------------------------------------------------------------- backbone
def mbconv(x, W, blk, bn_train, drop_scale=None, stats=None):
    p = blk["prefix"]
    inp = x
    if blk["expand"] != 1:
        x = conv2d(x, W[p + "expand_conv/kernel"])
        x = swish(batchnorm(x, W, p + "expand_bn", BN_EPS_BACKBONE, bn_train, stats))
    x = dwconv2d(x, W[p + "dwconv/depthwise_kernel"], blk["stride"])
    x = swish(batchnorm(x, W, p + "bn", BN_EPS_BACKBONE, bn_train, stats))
    se = x.mean(dim=(2, 3), keepdim=True)
    se = swish(conv2d(se, W[p + "se_reduce/kernel"], 1, W[p + "se_reduce/bias"]))
    se = torch.sigmoid(conv2d(se, W[p + "se_expand/kernel"], 1, W[p + "se_expand/bias"]))
    x = x * se
    x = conv2d(x, W[p + "project_conv/kernel"])
    x = batchnorm(x, W, p + "project_bn", BN_EPS_BACKBONE, bn_train, stats)
    if blk["skip"]:
        if drop_scale is not None and p in drop_scale:      # FixedDropout, noise (B,1,1,1)
            x = x * torch.as_tensor(drop_scale[p]).to(x.dtype).view(-1, 1, 1, 1)
        x = x + inp
    return x


def backbone(x, W, phi, bn_train=False, drop_scale=None, stats=None, force=None):
    blocks, taps = block_list(phi)
    x = conv2d(x, W["stem_conv/kernel"], 2)
    x = swish(batchnorm(x, W, "stem_bn", BN_EPS_BACKBONE, bn_train, stats))
    feats = []
    for i, blk in enumerate(blocks):
        x = forced(mbconv(x, W, blk, bn_train, drop_scale, stats), force, blk["prefix"] + "out")
        if i in taps:
            feats.append(x)
    return feats


# ---------------------------------------------------------------- BiFPN
def conv_block(x, W, name, k=1, s=1, bn_train=False, stats=None, force=None):          # model.py:71-90
    x = conv2d(x, W[name + "_conv/kernel"], s)
    return relu_forced(batchnorm(x, W, name + "_bn", BN_EPS_BIFPN, bn_train, stats), force, name)


def dw_block(x, W, name, bn_train=False, stats=None, force=None):                      # model.py:48-68
    x = dwconv2d(x, W[name + "_dconv/depthwise_kernel"], 1)
    return relu_forced(batchnorm(x, W, name + "_bn", BN_EPS_BIFPN, bn_train, stats), force, name)


def fuse(inputs, W, weighted, name, eps=1e-4):                             # layers.py:26-31
    if not weighted:
        x = inputs[0]
        for t in inputs[1:]:
            x = x + t
        return x
    w = torch.relu(torch.as_tensor(W[name + "/" + name]).to(inputs[0].dtype))
    x = w[0] * inputs[0]
    for i in range(1, len(inputs)):
        x = x + w[i] * inputs[i]
    return x / (w.sum() + eps)


def bifpn_layer(feats, W, i, weighted, bn_train=False, stats=None, force=None):
    kw = dict(bn_train=bn_train, stats=stats, force=force)
    pre = "BiFPN_%d_" % i
    if i == 0:
        _, _, C3, C4, C5 = feats
        P3 = conv_block(C3, W, pre + "P3", **kw)
        P4 = conv_block(C4, W, pre + "P4", **kw)
        P5 = conv_block(C5, W, pre + "P5", **kw)
        P6 = conv_block(C5, W, pre + "P6", 3, 2, **kw)
        P7 = conv_block(P6, W, pre + "P7", 3, 2, **kw)
    else:
        P3, P4, P5, P6, P7 = [conv_block(f, W, pre + "P%d" % (3 + j), **kw)
                              for j, f in enumerate(feats)]
    fn = lambda j: "w_bi_fpn_add" if (8 * i + j) == 0 else "w_bi_fpn_add_%d" % (8 * i + j)
    P6_td = dw_block(fuse([upsample2(P7), P6], W, weighted, fn(0)), W, pre + "U_P6", **kw)
    P5_td = dw_block(fuse([upsample2(P6_td), P5], W, weighted, fn(1)), W, pre + "U_P5", **kw)
    P4_td = dw_block(fuse([upsample2(P5_td), P4], W, weighted, fn(2)), W, pre + "U_P4", **kw)
    P3_o = dw_block(fuse([upsample2(P4_td), P3], W, weighted, fn(3)), W, pre + "U_P3", **kw)
    P4_o = dw_block(fuse([maxpool2(P3_o), P4_td, P4], W, weighted, fn(4)), W, pre + "D_P4", **kw)
    P5_o = dw_block(fuse([maxpool2(P4_o), P5_td, P5], W, weighted, fn(5)), W, pre + "D_P5", **kw)
    P6_o = dw_block(fuse([maxpool2(P5_o), P6_td, P6], W, weighted, fn(6)), W, pre + "D_P6", **kw)
    P7_o = dw_block(fuse([maxpool2(P6_o), P7], W, weighted, fn(7)), W, pre + "D_P7", **kw)
    return [P3_o, P4_o, P5_o, P6_o, P7_o]


# ---------------------------------------------------------------- heads
def head(x, W, scope, trunk_fmt, final_name, depth, force=None, level=0):
    for i in range(depth):
        n = scope + "/" + trunk_fmt % i
        x = relu_forced(conv2d(x, W[n + "/kernel"], 1, W[n + "/bias"]), force, "%s_%d_l%d" % (scope, i, level))
    n = scope + "/" + final_name
    return conv2d(x, W[n + "/kernel"], 1, W[n + "/bias"])


def forward(W, images, phi, num_classes, weighted_bifpn=False, dtype=torch.float32,
            bn_train_bifpn=False, bn_train_backbone=False, drop_scale=None, taps=None,
            stats=None, force=None):
    """images: (B,S,S,3) array/tensor.  W values may be numpy arrays or torch tensors
    (leaf tensors with requires_grad for the training oracle).  `force`: see forced().
    Returns (regression (B,N,4), classification (B,N,C)) torch tensors; `taps`, if a
    dict, receives NHWC copies of C1..C5 and every BiFPN layer's outputs."""
    x = torch.as_tensor(images).to(dtype).permute(0, 3, 1, 2)
    feats = backbone(x, W, phi, bn_train_backbone, drop_scale, stats, force)
    if taps is not None:
        for j, f in enumerate(feats):
            taps["C%d" % (j + 1)] = f.permute(0, 2, 3, 1).detach()
    depth = 3 + phi // 3
    for i in range(2 + phi):
        feats = bifpn_layer(feats, W, i, weighted_bifpn, bn_train_bifpn, stats, force)
        if taps is not None:
            for j, f in enumerate(feats):
                taps["BiFPN_%d_P%d" % (i, j + 3)] = f.permute(0, 2, 3, 1).detach()
    B = x.shape[0]
    regs, clss = [], []
    for l, f in enumerate(feats):
        r = head(f, W, "box_head", "regress_head_conv_%d", "regress_head_conv_final", depth, force, l)
        regs.append(r.permute(0, 2, 3, 1).reshape(B, -1, 4))
        c = head(f, W, "class_head", "class_head_%d", "pyramid_classification", depth, force, l)
        clss.append(torch.sigmoid(c.permute(0, 2, 3, 1).reshape(B, -1, num_classes)))
    return torch.cat(regs, 1), torch.cat(clss, 1)


def keras_layer_count(phi, drop_connect_rate=0.2):
    """Number of Keras layers input..block7*_project_bn/add:
    must reproduce train_tpu.py:24 EFFICIENTNET_DEPTHS."""
    blocks, _ = block_list(phi)
    n = 1 + 3
    for b in blocks:
        n += (3 if b["expand"] != 1 else 0) + 3 + 5 + 2
        if b["skip"]:
            rate = drop_connect_rate * b["num"] / 16.0   # num_blocks_total uses UNSCALED repeats
            n += 1 + (1 if rate > 0 else 0)
    return n
