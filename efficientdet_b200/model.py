"""efficientdet(phi, ...) builder -- mirror of the reference's model.py:356-452 public surface.

Same signature and return convention (`model` or `(model, prediction_model)`); the returned
objects expose the slice of the Keras Model API the reference's callers use
(predict_on_batch, load_weights(by_name=True), layers[i].trainable, compile, fit /
train_on_batch, summary, optimizer) and execute on the library's CUDA kernels through
engine.Plan.  Extra keyword-only options of this build (popped before **bbkwargs reach the
backbone builder): dtype ('fp32' accuracy mode | 'bf16' speed mode), image_size (override of
model.py:29), seed (numpy default_rng seed of the random initialisation), device,
tensor_cores (bf16 mode: False forces the SIMT convolution kernels).
"""
import collections
import math

import numpy as np
import torch

from . import _lib, engine
from ._lib import BF16, F32
from .efficientnet import (EfficientNetB0, EfficientNetB1, EfficientNetB2, EfficientNetB3,
                           EfficientNetB4, EfficientNetB5, EfficientNetB6)
from .initializers import PriorProbability

w_bifpns = [64, 88, 112, 160, 224, 288, 384]
image_sizes = [512, 640, 768, 896, 1024, 1280, 1408]
backbones = [EfficientNetB0, EfficientNetB1, EfficientNetB2, EfficientNetB3, EfficientNetB4,
             EfficientNetB5, EfficientNetB6]
batchnorm_config = {"momentum": 0.997, "epsilon": 1e-4}     # model.py:42-45
EFFICIENTNET_DEPTHS = [227, 329, 329, 374, 464, 566, 656]  # train_tpu.py:24


def _fuse_name(k):
    return "w_bi_fpn_add" if k == 0 else "w_bi_fpn_add_%d" % k


def _is_u8(images):
    return (images.dtype == torch.uint8) if isinstance(images, torch.Tensor) else \
        (getattr(images, "dtype", None) == np.uint8)


class Network:
    """Weights (Keras names, fp32 masters on the device) + derived tensors + plans."""

    def __init__(self, phi, num_classes, weighted_bifpn, freeze_bn, backbone, image_size, dtype,
                 seed, device, tensor_cores=True):
        self.phi, self.num_classes = phi, int(num_classes)
        self.weighted_bifpn, self.freeze_bn = bool(weighted_bifpn), bool(freeze_bn)
        self.backbone = backbone
        self.image_size = int(image_size)
        if self.image_size % 128 != 0:
            raise ValueError("image size must be a multiple of 128 (5 pyramid levels, stride 128)")
        self.dtype = {"fp32": F32, "float32": F32, "bf16": BF16, "bfloat16": BF16}[dtype]
        self.device = device
        self.w_bifpn = w_bifpns[phi]
        self.d_bifpn = 2 + phi
        self.head_depth = 3 + int(phi / 3)
        self.use_tensor_cores = bool(tensor_cores)   # bf16 mode: tcgen05 convolutions; fp32 mode: split-bf16 (3 terms)
        self.weights = collections.OrderedDict()
        self.bn_layers = collections.OrderedDict()     # bn layer name -> (C, eps)
        self.folded = {}
        self.drop_scale = {}
        self.frozen_layers = set()
        self.plans = {}
        self._staged = collections.OrderedDict()
        self.seed = seed
        self._init_weights(seed)
        self._materialize()
        self.refresh()

    # ---------------------------------------------------------------- initialisation
    def _put(self, name, arr):
        self._staged[name] = np.ascontiguousarray(arr, np.float32)

    def _materialize(self):
        """All weights live in ONE flat fp32 device buffer (16-byte aligned slots, creation
        order: backbone, BiFPN, heads) so that the gradient all-reduce and the SGD update run
        over contiguous ranges; `weights[name]` are views into it."""
        offs, total = collections.OrderedDict(), 0
        for name, a in self._staged.items():
            offs[name] = total
            total += (a.size + 3) // 4 * 4
        host = np.zeros(total, np.float32)
        for name, a in self._staged.items():
            host[offs[name]:offs[name] + a.size] = a.reshape(-1)
        self.flat = torch.from_numpy(host).to(self.device)
        self.offsets = offs
        for name, a in self._staged.items():
            self.weights[name] = self.flat[offs[name]:offs[name] + a.size].view(a.shape)
        first_bifpn = next(k for k in offs if k.startswith("BiFPN_"))
        self.backbone_end = offs[first_bifpn]      # flat index where BiFPN + head params start
        self._staged = None

    def _bn(self, name, C, eps):
        self._put(name + "/gamma", np.ones(C))
        self._put(name + "/beta", np.zeros(C))
        self._put(name + "/moving_mean", np.zeros(C))
        self._put(name + "/moving_variance", np.ones(C))
        self.bn_layers[name] = (C, eps)

    def _init_weights(self, seed):
        rng = np.random.default_rng(seed)

        def vs_normal(shape, fan_out):      # efficientnet.py:116-129 CONV_KERNEL_INITIALIZER
            return rng.standard_normal(shape) * math.sqrt(2.0 / fan_out)

        def glorot(shape, fan_in, fan_out):  # keras default kernel initializer (model.py:49-55,72-79)
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            return rng.uniform(-lim, lim, shape)
        bb = self.backbone
        eps = engine.BN_EPS_BACKBONE
        self._put("stem_conv/kernel", vs_normal((3, 3, 3, bb.stem_filters), 9 * bb.stem_filters))
        self._bn("stem_bn", bb.stem_filters, eps)
        for b in bb.blocks:
            p, cin, cmid, cout, k = b.prefix, b.input_filters, b.mid_filters, b.output_filters, b.kernel_size
            if b.expand_ratio != 1:
                self._put(p + "expand_conv/kernel", vs_normal((1, 1, cin, cmid), cmid))
                self._bn(p + "expand_bn", cmid, eps)
            self._put(p + "dwconv/depthwise_kernel", vs_normal((k, k, cmid, 1), k * k))
            self._bn(p + "bn", cmid, eps)
            self._put(p + "se_reduce/kernel", vs_normal((1, 1, cmid, b.se_filters), b.se_filters))
            self._put(p + "se_reduce/bias", np.zeros(b.se_filters))
            self._put(p + "se_expand/kernel", vs_normal((1, 1, b.se_filters, cmid), cmid))
            self._put(p + "se_expand/bias", np.zeros(cmid))
            self._put(p + "project_conv/kernel", vs_normal((1, 1, cmid, cout), cout))
            self._bn(p + "project_bn", cout, eps)
        feat_ch = [bb.blocks[i].output_filters for i in bb.feature_after]
        W, eps = self.w_bifpn, batchnorm_config["epsilon"]
        for i in range(self.d_bifpn):
            pre = "BiFPN_%d_" % i
            for l in range(3, 8):
                if i == 0:
                    cin = feat_ch[l - 1] if l <= 5 else (feat_ch[4] if l == 6 else W)
                    k = 1 if l <= 5 else 3
                else:
                    cin, k = W, 1
                self._put(pre + "P%d_conv/kernel" % l, glorot((k, k, cin, W), k * k * cin, k * k * W))
                self._bn(pre + "P%d_bn" % l, W, eps)
            for j, nm in enumerate(["U_P6", "U_P5", "U_P4", "U_P3", "D_P4", "D_P5", "D_P6", "D_P7"]):
                self._put(pre + nm + "_dconv/depthwise_kernel", glorot((3, 3, W, 1), 9 * W, 9))
                self._bn(pre + nm + "_bn", W, eps)
                if self.weighted_bifpn:
                    n_in = 3 if nm in ("D_P4", "D_P5", "D_P6") else 2
                    fn = _fuse_name(8 * i + j)
                    self._put(fn + "/" + fn, np.full((n_in,), 1.0 / n_in))
        A, C = 9, self.num_classes
        for i in range(self.head_depth):
            self._put("box_head/regress_head_conv_%d/kernel" % i, rng.normal(0, 0.01, (3, 3, W, W)))
            self._put("box_head/regress_head_conv_%d/bias" % i, np.zeros(W))
        self._put("box_head/regress_head_conv_final/kernel", rng.normal(0, 0.01, (3, 3, W, A * 4)))
        self._put("box_head/regress_head_conv_final/bias", np.zeros(A * 4))
        for i in range(self.head_depth):
            self._put("class_head/class_head_%d/kernel" % i, rng.normal(0, 0.01, (3, 3, W, W)))
            self._put("class_head/class_head_%d/bias" % i, np.zeros(W))
        self._put("class_head/pyramid_classification/kernel", rng.normal(0, 0.01, (3, 3, W, A * C)))
        self._put("class_head/pyramid_classification/bias", PriorProbability(0.01)((A * C,)))

    # ---------------------------------------------------------------- derived tensors
    def refresh(self):
        """Re-derive folded BatchNorm scale/shift after the weights changed."""
        st = _lib.stream_ptr(self.device)
        for name, (C, eps) in self.bn_layers.items():
            if name not in self.folded:
                self.folded[name] = (torch.empty(C, dtype=torch.float32, device=self.device),
                                     torch.empty(C, dtype=torch.float32, device=self.device))
            sc, sh = self.folded[name]
            _lib.call("effdet_bn_fold", self.weights[name + "/gamma"].data_ptr(),
                      self.weights[name + "/beta"].data_ptr(),
                      self.weights[name + "/moving_mean"].data_ptr(),
                      self.weights[name + "/moving_variance"].data_ptr(), float(eps),
                      sc.data_ptr(), sh.data_ptr(), C, st)
        for k in getattr(self, "_panels", {}):
            self._build_panel(k)

    # ---------------------------------------------------------------- training state
    def ensure_grad_buffers(self):
        if not hasattr(self, "grad_flat"):
            self.grad_flat = torch.zeros_like(self.flat)
            self.velocity = torch.zeros_like(self.flat)
            self.grads = {k: self.grad_flat[o:o + self.weights[k].numel()].view(self.weights[k].shape)
                          for k, o in self.offsets.items()}
            self._consts = {}

    def const_ones(self, C):
        k = ("ones", C)
        if k not in self._consts:
            self._consts[k] = torch.ones(C, dtype=torch.float32, device=self.device)
        return self._consts[k]

    def const_zeros(self, C):
        k = ("zeros", C)
        if k not in self._consts:
            self._consts[k] = torch.zeros(C, dtype=torch.float32, device=self.device)
        return self._consts[k]

    def invalidate(self):
        """Weights changed (optimizer step): inference-mode folded BN must be re-derived."""
        self._dirty = True

    def static_panel(self, key, taps, cin, cout, mode):
        """bf16 weight panel for the tcgen05 convolution, cached per (weight, mode)."""
        if not hasattr(self, "_panels"):
            self._panels = {}
        k = (key, mode)
        if k not in self._panels:
            if mode == "split":
                n = _lib.load().effdet_conv_weight_panel_split_elems(taps, cin, cout)
            else:
                n = _lib.load().effdet_conv_weight_panel_elems(taps, cin if mode == 0 else cout,
                                                               cout if mode == 0 else cin)
            t = torch.empty(n, dtype=torch.bfloat16, device=self.device)
            self._panels[k] = (t, taps, cin, cout)
            self._build_panel(k)
        return self._panels[k][0]

    def _build_panel(self, k):
        t, taps, cin, cout = self._panels[k]
        if k[1] == "split":       # fp32 accuracy mode on the tensor cores: [Whi | Whi | Wlo]
            _lib.call("effdet_conv_weight_panel_split", self.weights[k[0]].data_ptr(), t.data_ptr(), taps, cin,
                      cout, None, 0, _lib.stream_ptr(self.device))
            return
        _lib.call("effdet_conv_weight_panel", self.weights[k[0]].data_ptr(), t.data_ptr(), taps, cin, cout,
                  k[1], None, 0, _lib.stream_ptr(self.device))

    def normalization_lut(self):
        """(3, 256) f32 on the device: lut[c][v] = ((v / 255) - mean_c) / std_c evaluated in float32 the way the
        reference does (train_tpu.py:130-140, generators/common.py:418-429) -- see utils.preprocess."""
        if getattr(self, "_norm_lut", None) is None:
            from .utils.preprocess import normalization_lut
            self._norm_lut = torch.from_numpy(normalization_lut()).to(self.device)
        return self._norm_lut

    def plan(self, batch, **kw):
        if getattr(self, "_dirty", False):
            self.refresh()
            self._dirty = False
        key = (int(batch), tuple(sorted(kw.items())))
        if key not in self.plans:
            self.plans[key] = engine.Plan(self, batch, **kw)
        return self.plans[key]

    # ---------------------------------------------------------------- weight exchange
    def get_weights_dict(self):
        return {k: v.detach().cpu().numpy() for k, v in self.weights.items()}

    def set_weights_dict(self, d, strict=False):
        n = 0
        for k, v in d.items():
            k = k[:-2] if k.endswith(":0") else k
            if k in self.weights:
                t = self.weights[k]
                a = np.asarray(v, np.float32)
                if tuple(a.shape) != tuple(t.shape):
                    raise ValueError("shape mismatch for %s: %s vs %s" % (k, a.shape, tuple(t.shape)))
                t.copy_(torch.from_numpy(a))
                n += 1
            elif strict:
                raise KeyError(k)
        self.refresh()
        return n


class LayerRef:
    """Handle returned by model.layers[i]: name + trainable flag (train.py:338-339)."""

    def __init__(self, net, name):
        self._net, self.name = net, name

    @property
    def trainable(self):
        return self.name not in self._net.frozen_layers

    @trainable.setter
    def trainable(self, value):
        if value:
            self._net.frozen_layers.discard(self.name)
        else:
            self._net.frozen_layers.add(self.name)


class Model:
    """The slice of keras.Model the reference's callers use."""

    def __init__(self, net, name, outputs, anchors=None, score_threshold=0.01, takes_anchors=False):
        self.net, self.name, self._outputs = net, name, outputs
        self._anchors = anchors            # baked (1,N,4) float32 CUDA tensor or None
        self._takes_anchors = takes_anchors
        self.score_threshold = score_threshold
        self.output_names = {"train": ["regression", "classification"],
                             "boxes": ["clipped_boxes", "classification"],
                             "detections": ["filtered_detections"] * 3}[outputs]
        self.optimizer = None
        self.loss = None
        self._trainer = None
        self._pinned = {}
        names = ["input_1"] + net.backbone.keras_layer_names()
        for i in range(net.d_bifpn):
            pre = "BiFPN_%d_" % i
            for l in range(3, 8):
                names += [pre + "P%d_conv" % l, pre + "P%d_bn" % l, pre + "P%d_relu" % l]
            for j, nm in enumerate(["U_P6", "U_P5", "U_P4", "U_P3", "D_P4", "D_P5", "D_P6", "D_P7"]):
                if net.weighted_bifpn:
                    names.append(_fuse_name(8 * i + j))
                names += [pre + nm + "_dconv", pre + nm + "_bn", pre + nm + "_relu"]
        names += ["box_head", "class_head", "regression", "classification"]
        if outputs != "train":
            names += ["boxes", "clipped_boxes"] + (["filtered_detections"] if outputs == "detections" else [])
        self.layers = [LayerRef(net, n) for n in names]

    @property
    def backbone_depth(self):
        """Index one past the last backbone layer in `layers` (== EFFICIENTNET_DEPTHS[phi] for the
        default drop_connect_rate; smaller when drop_connect_rate=0 removes the `*_drop` layers)."""
        return 1 + len(self.net.backbone.keras_layer_names())

    def freeze_backbone(self):
        """train_tpu.py:272-274 / train.py:336-339 `--freeze-backbone`."""
        for i in range(1, self.backbone_depth):
            self.layers[i].trainable = False

    # ---------------------------------------------------------------- inference
    def _stage(self, images):
        """host numpy -> pinned staging buffer -> device (async)."""
        u8 = _is_u8(images)
        if isinstance(images, torch.Tensor):
            return images.to(self.net.device, torch.uint8 if u8 else torch.float32)
        a = np.ascontiguousarray(images, np.uint8 if u8 else np.float32)
        key = (a.shape, u8)
        if key not in self._pinned:
            self._pinned[key] = [torch.empty(a.shape, dtype=torch.uint8 if u8 else torch.float32).pin_memory(),
                                 None]
        slot = self._pinned[key]
        if slot[1] is not None:
            slot[1].synchronize()       # the previous async copy out of this buffer must have finished
        slot[0].numpy()[...] = a
        d = slot[0].to(self.net.device, non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record(torch.cuda.current_stream(self.net.device))
        return d

    def predict_on_batch_device(self, x, defer_status=False):
        """Same as predict_on_batch but returns CUDA tensors (no device->host copy)."""
        if isinstance(x, (list, tuple)):
            images = x[0]
            anchors = x[1] if len(x) > 1 else None
        else:
            images, anchors = x, None
        net = self.net
        S = net.image_size
        if tuple(images.shape[1:]) != (S, S, 3):
            raise ValueError("expected images of shape (B, %d, %d, 3), got %s" % (S, S, tuple(images.shape)))
        B = int(images.shape[0])
        # uint8 input = the raw letterboxed RGB image (utils.preprocess.preprocess_image); normalize_image runs
        # inside the stem kernel
        plan = net.plan(B, u8_input=True) if _is_u8(images) else net.plan(B)
        reg, cls = plan.forward(self._stage(images))
        if self._outputs == "train":
            return [reg, cls]
        if self._takes_anchors:
            if anchors is None:
                raise ValueError("this model takes [images, anchors] (model.py:421-427)")
            a = torch.as_tensor(anchors).to(net.device, torch.float32).contiguous()
        else:
            a = self._anchors
        if a.dim() != 3 or a.shape[1] != plan.N or a.shape[0] not in (1, B):
            raise ValueError("anchors must be (1 or B, %d, 4)" % plan.N)
        boxes = torch.empty((B, plan.N, 4), dtype=torch.float32, device=net.device)
        f4 = _lib.c_float * 4
        _lib.call("effdet_regress_clip_boxes", a.data_ptr(), int(a.shape[0] == B and B > 1),
                  reg.data_ptr(), f4(0, 0, 0, 0), f4(.2, .2, .2, .2), B, plan.N, float(S), float(S),
                  boxes.data_ptr(), _lib.stream_ptr(net.device))
        if self._outputs == "boxes":
            return [boxes, cls]
        if defer_status:        # predict_generator: no host synchronisation here; the caller checks `status` later
            from .FilterDetections import _capacity, _launch
            cap, _ = _capacity(B, plan.N, int(cls.shape[2]), int(cls.shape[2]))
            ob, os_, ol, _, status = _launch(boxes, cls, True, self.score_threshold, 300, 0.5, True, cap)
            return [ob, os_, ol], status
        from .FilterDetections import _run
        return list(_run(boxes, cls, True, self.score_threshold, 300, 0.5, True))

    def predict_on_batch(self, x):
        outs = self.predict_on_batch_device(x)
        return [o.cpu().numpy() for o in outs]

    def predict_generator(self, batches):
        """keras Model.predict_generator over an iterable of HOST image batches ((B,S,S,3) float32 -- or raw
        uint8 -- torch tensors, ideally pinned): the host->device copy of batch i+1 runs on a copy stream while batch i
        is computed (what the reference gets from keras' generator queue / tf.data prefetch).  Yields
        the outputs of predict_on_batch (numpy) per batch."""
        dev = self.net.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
            self._slots = {}
        cs = self._copy_stream

        def stage(slot, xb):
            xb = torch.as_tensor(xb)
            dt = torch.uint8 if xb.dtype == torch.uint8 else torch.float32
            key = (slot, tuple(xb.shape), dt)
            if key not in self._slots:
                self._slots[key] = torch.empty(xb.shape, dtype=dt, device=dev)
            d = self._slots[key]
            with torch.cuda.stream(cs):
                d.copy_(xb, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            return d, ev
        # Outputs (and the tail's overflow status) of batch i are copied to pinned host buffers behind batch i on the
        # main stream and waited for one iteration later, when batch i+1 is already enqueued -- the GPU no longer
        # idles while the host launches the next batch (17 % of a 0.74 ms batch-1 step, 6 % at D2 / batch 64).  A
        # staging slot is rewritten only after the batch that used it has consumed it (`consumed` events).  If the
        # candidate workspace overflowed (rare: the capacity is learned), that batch is redone synchronously.
        detect = self._outputs not in ("train", "boxes")
        consumed = [None, None]
        it = iter(batches)
        first = next(it, None)
        nxt = stage(0, first) if first is not None else None
        host_cur = first
        i = 0
        pending = None

        def finish(p):
            host_outs, status_h, lev, host_batch = p
            lev.synchronize()
            if status_h is not None and int(status_h[0]) != 0:
                from .FilterDetections import _grow_capacity
                B, N, C = int(host_outs[0].shape[0]), self.net.plan(int(host_outs[0].shape[0])).N, self.net.num_classes
                _grow_capacity(B, N, C, C, status_h)
                return self.predict_on_batch([host_batch])
            return [o.numpy().copy() for o in host_outs]
        while nxt is not None:
            cur, ev = nxt
            b = next(it, None)
            if b is not None:
                slot = (i + 1) % 2
                if consumed[slot] is not None:
                    cs.wait_event(consumed[slot])
                nxt = stage(slot, b)
            else:
                nxt = None
            main.wait_event(ev)
            if detect:
                outs, status = self.predict_on_batch_device([cur], defer_status=True)
            else:
                outs, status = self.predict_on_batch_device([cur]), None
            cev = torch.cuda.Event()
            cev.record(main)
            consumed[i % 2] = cev
            key = ("out", i % 2, tuple(tuple(o.shape) for o in outs))
            if key not in self._slots:
                self._slots[key] = ([torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs],
                                    torch.empty(4, dtype=torch.int32).pin_memory())
            host_outs, status_h = self._slots[key]
            for h, o in zip(host_outs, outs):
                h.copy_(o, non_blocking=True)
            if status is not None:
                status_h.copy_(status, non_blocking=True)
            lev = torch.cuda.Event()
            lev.record(main)
            if pending is not None:
                yield finish(pending)
            pending = (host_outs, status_h if status is not None else None, lev, host_cur)
            host_cur = b
            i += 1
        if pending is not None:
            yield finish(pending)

    def predict(self, x, batch_size=32, **kw):
        images = x[0] if isinstance(x, (list, tuple)) else x
        rest = list(x[1:]) if isinstance(x, (list, tuple)) else []
        outs = []
        for i in range(0, len(images), batch_size):
            outs.append(self.predict_on_batch([images[i:i + batch_size]] + rest))
        return [np.concatenate([o[j] for o in outs], 0) for j in range(len(outs[0]))]

    # ---------------------------------------------------------------- weights
    def get_weights_dict(self):
        return self.net.get_weights_dict()

    def set_weights_dict(self, d, strict=False):
        return self.net.set_weights_dict(d, strict)

    def save_weights(self, path):
        """`.h5` / `.hdf5`: Keras weight-file layout (one group per layer, `layer_names` / `weight_names`
        attributes; utils/hdf5.py) so that the reference's `load_weights(path, by_name=True)` (train.py:329-332)
        can read it; anything else: `.npz` keyed '<layer>/<weight>'."""
        if str(path).endswith((".h5", ".hdf5")):
            from .utils import hdf5
            hdf5.save_keras_weights(path, self.net.get_weights_dict(),
                                    layer_order=list(dict.fromkeys(k.split("/")[0] for k in self.net.weights)))
            return
        np.savez(path, **self.net.get_weights_dict())

    def load_weights(self, path, by_name=True, skip_mismatch=False):
        """Keras `.h5` weight files (the reference's exchange format, train.py:329-332, utils/train.py:10-35;
        read by the built-in HDF5 subset reader utils/hdf5.py -- h5py is not needed) or the `.npz` written by
        save_weights; keys = '<keras layer>/<weight>[:0]'.  by_name=True ignores unknown / missing names like
        Keras does; skip_mismatch drops weights whose shape differs."""
        if str(path).endswith((".h5", ".hdf5")):
            from .utils import hdf5
            d = hdf5.load_keras_weights(path)
        else:
            d = {(k[:-2] if k.endswith(":0") else k): v for k, v in dict(np.load(path)).items()}
        if skip_mismatch:
            d = {k: v for k, v in d.items() if k in self.net.weights and
                 tuple(v.shape) == tuple(self.net.weights[k].shape)}
        return self.net.set_weights_dict(d, strict=not by_name)

    def summary(self, print_fn=print):
        print_fn('Model: "%s"' % self.name)
        tot = 0
        for k, v in self.net.weights.items():
            print_fn("%-60s %s" % (k, tuple(v.shape)))
            tot += v.numel()
        print_fn("Total params: {:,}".format(tot))

    def count_params(self):
        return sum(v.numel() for v in self.net.weights.values())

    # ---------------------------------------------------------------- training (train.py)
    def compile(self, optimizer=None, loss=None, **kw):
        from . import train
        self.loss = loss
        self._trainer = train.Trainer(self, optimizer, loss)
        self.optimizer = self._trainer.opt      # the default SGD when none was given (callbacks read .optimizer.lr)

    def train_on_batch(self, x, y):
        if self._trainer is None:
            raise RuntimeError("compile() the model before training")
        return self._trainer.step(x, y)

    def fit(self, x=None, y=None, epochs=1, steps_per_epoch=None, batch_size=None, verbose=0, callbacks=None,
            initial_epoch=0, **kw):
        """x: iterable of (images, (regression_targets, class_targets)) batches, or arrays.
        callbacks: keras-style objects with optional set_model / on_epoch_begin / on_epoch_end
        (utils.lr_schedule.LearningRateScheduler, utils.train.CheckpointSaver; train.py:377-388)."""
        callbacks = list(callbacks or [])
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
        hist = []
        for epoch in range(initial_epoch, epochs):
            for cb in callbacks:
                if hasattr(cb, "on_epoch_begin"):
                    cb.on_epoch_begin(epoch, {})
            it = iter(x) if y is None else iter([(x, y)])
            last = None
            for step, (xb, yb) in enumerate(it):
                if steps_per_epoch is not None and step >= steps_per_epoch:
                    break
                last = self.train_on_batch(xb, yb)
                hist.append(last)
            logs = {} if last is None else {"loss": last[0], "regression_loss": last[1],
                                            "classification_loss": last[2]}
            for cb in callbacks:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(epoch, logs)
        return hist


def efficientdet(phi, num_classes=20, weighted_bifpn=False, freeze_bn=False, score_threshold=0.01,
                 no_filter=False, anchors=None, just_training_model=False, **bbkwargs):
    """Builds EfficientDet-D{phi}.  Returns `model` if just_training_model else
    `(model, prediction_model)` sharing weights (model.py:356-452)."""
    assert phi in range(7)
    dtype = bbkwargs.pop("dtype", "fp32")
    image_size = bbkwargs.pop("image_size", None) or image_sizes[phi]
    seed = bbkwargs.pop("seed", 2024)
    device = bbkwargs.pop("device", None)
    tensor_cores = bbkwargs.pop("tensor_cores", True)
    if device is None:
        from ._tensor import device as _dev
        device = _dev()
    backbone = backbones[phi](input_tensor=None, freeze_bn=freeze_bn, **bbkwargs)
    net = Network(phi, num_classes, weighted_bifpn, freeze_bn, backbone, image_size, dtype, seed,
                  device, tensor_cores)
    model = Model(net, "efficientdet", "train")
    if just_training_model:
        return model
    baked = None
    if anchors is not None:
        arr = np.expand_dims(np.asarray(anchors), axis=0).astype("float32")
        baked = torch.from_numpy(arr).to(device)
        net.weights["boxes/anchor_boxes_baked"] = baked
    prediction_model = Model(net, "efficientdet_p", "boxes" if no_filter else "detections",
                             anchors=baked, score_threshold=score_threshold,
                             takes_anchors=anchors is None)
    return model, prediction_model
