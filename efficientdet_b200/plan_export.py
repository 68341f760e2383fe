"""Compiled training plans: the launch list of one optimizer step written to a file that the C library replays
without Python (`effdet_replay_load / effdet_replay_region / effdet_replay_step`, csrc/replay.cu).

The lowering of the training step (target assignment -> forward -> losses -> backward -> SGD, train_tpu.py:249-346)
lives in train.py; this module serialises its RESULT -- for every launch the C entry point's name and its arguments,
with every device pointer rewritten as (region, offset) -- plus the initial contents of the regions that hold state
(weights, optimizer velocity, folded BatchNorms, static weight panels, anchors, constants).  A host in any language
then trains with three calls; the launches, their order and their arguments are exactly the ones the Python plan
issues, so the results are bit-identical (tests/test_gpu_replay.py).

File layout (little endian):
  "EFDPLAN1"
  u32 n_regions; per region: u64 bytes, u8 has_content, u16 name_len, name, [content]
  u32 n_ops;     per op: u8 lane, u16 n_waits, n_waits x u32 (launches on other lanes to wait for: the multi-lane
                 capture order of engine.Plan.lane_schedule), u16 name_len, name, u16 n_args, per arg: u8 kind, payload
     kind 0 i64 | 1 f64 | 2 device pointer (u32 region, u64 offset) | 3 NULL | 4 host blob (u32 len, bytes,
     u16 n_reloc, per reloc: u32 byte offset in the blob, u32 region, u64 offset) | 5 the step's learning rate (f64)
"""
import bisect
import ctypes
import gc
import struct

import numpy as np
import torch

from . import _lib

I64, F64, PTR, NULL, BLOB, LR = range(6)
LR_MARK = object()          # placeholder argument: replaced by the learning rate effdet_replay_step() is given


class _Regions:
    """Device allocations referenced by the plan: every live CUDA tensor's storage is a candidate."""

    def __init__(self, device):
        seen = {}
        for o in gc.get_objects():
            try:
                if isinstance(o, torch.Tensor) and o.is_cuda and o.device == device:
                    st = o.untyped_storage()
                    if st.nbytes() > 0 and st.nbytes() >= seen.get(st.data_ptr(), (0, None))[0]:
                        seen[st.data_ptr()] = (st.nbytes(), st)
            except Exception:
                continue
        self.bases = sorted(seen)
        self.sizes = [seen[b][0] for b in self.bases]
        self.storages = {b: seen[b][1] for b in self.bases}
        self.used = {}          # base -> region index
        self.names = {}         # region index -> name
        self.order = []

    def locate(self, ptr):
        i = bisect.bisect_right(self.bases, ptr) - 1
        if i < 0 or ptr >= self.bases[i] + self.sizes[i]:
            raise ValueError("pointer 0x%x is not inside any live CUDA tensor" % ptr)
        base = self.bases[i]
        if base not in self.used:
            self.used[base] = len(self.order)
            self.order.append((base, self.sizes[i]))
        return self.used[base], ptr - base

    def name(self, tensor, name):
        r, off = self.locate(tensor.data_ptr())
        if off != 0:
            raise ValueError("named region %s must start its allocation" % name)
        self.names[r] = name
        return r


def _is_ptr_type(t):
    return t is ctypes.c_void_p or (isinstance(t, type) and issubclass(t, ctypes.Array))


def _blob(obj, regions):
    """ctypes struct / array -> (bytes, relocations): pointer-typed fields become (offset, region, offset)."""
    raw = bytes(memoryview(obj).cast("B"))
    rel = []

    def walk(o, base):
        if isinstance(o, ctypes.Structure):
            for fname, ftype in o._fields_:
                off = base + getattr(type(o), fname).offset
                if ftype is ctypes.c_void_p:
                    v = getattr(o, fname)
                    if v:
                        rel.append((off,) + regions.locate(int(v)))
                elif isinstance(ftype, type) and issubclass(ftype, ctypes.Array) and ftype._type_ is ctypes.c_void_p:
                    arr = getattr(o, fname)
                    for k in range(len(arr)):
                        if arr[k]:
                            rel.append((off + 8 * k,) + regions.locate(int(arr[k])))
        elif isinstance(o, ctypes.Array) and o._type_ is ctypes.c_void_p:
            for k in range(len(o)):
                if o[k]:
                    rel.append((base + 8 * k,) + regions.locate(int(o[k])))
    walk(obj, 0)
    return raw, rel


def _encode_call(name, args, regions):
    sig = _lib._SIGNATURES[name]
    if len(args) != len(sig) - 1:          # the stream is the last parameter of every launch entry point
        raise ValueError("%s: %d arguments recorded, signature has %d + stream" % (name, len(args), len(sig) - 1))
    out = [struct.pack("<H", len(name)), name.encode(), struct.pack("<H", len(args))]
    for a, t in zip(args, sig):
        if a is LR_MARK:
            out.append(struct.pack("<B", LR))
        elif a is None:
            out.append(struct.pack("<B", NULL))
        elif type(a).__name__ == "CArgObject":                      # ctypes.byref(struct)
            raw, rel = _blob(a._obj, regions)
            out.append(struct.pack("<BI", BLOB, len(raw)) + raw + struct.pack("<H", len(rel)) +
                       b"".join(struct.pack("<IIQ", *r) for r in rel))
        elif isinstance(a, (ctypes.Array, ctypes.Structure)):
            raw, rel = _blob(a, regions)
            out.append(struct.pack("<BI", BLOB, len(raw)) + raw + struct.pack("<H", len(rel)) +
                       b"".join(struct.pack("<IIQ", *r) for r in rel))
        elif _is_ptr_type(t):
            r, off = regions.locate(int(a))
            out.append(struct.pack("<BIQ", PTR, r, off))
        elif t in (ctypes.c_float, ctypes.c_double):
            out.append(struct.pack("<Bd", F64, float(a)))
        else:
            out.append(struct.pack("<Bq", I64, int(a)))
    return b"".join(out)


def export_train_plan(model, batch, path, kmax=16, u8_input=False):
    """Writes the training step of `model` (compiled: model.compile(optimizer=SGD(...))) for per-replica batch
    `batch` to `path`.  kmax: annotation slots per image of the gt_* input regions.
    Named regions of the file: images, anchors, gt_boxes (B,kmax,4) f64, gt_labels (B,kmax) i32, gt_counts (B,) i32,
    image_hw (B,2) f64, losses (8,) f32 [focal, smooth-L1, ...], weights (the flat fp32 parameter buffer)."""
    from .utils.anchors import anchors_for_shape
    tr, net = model._trainer, model.net
    if tr._frozen_keys():
        raise NotImplementedError("layers.trainable = False outside the backbone is applied on the host side "
                                  "(Trainer._reduce_and_update); not part of an exported plan")
    dev = net.device
    plan = tr.plan(int(batch), False, u8=u8_input)
    if plan.graph is None:
        # lazy kernel attributes / allocations happen on the first run; it is a real forward + backward, so the
        # BatchNorm moving statistics and the stochastic-depth step counter it advanced are put back
        snap = net.flat.clone()
        step = plan.drop_step.clone() if plan.drop_blocks else None
        plan.run()
        torch.cuda.synchronize(dev)
        net.flat.copy_(snap)
        if step is not None:
            plan.drop_step.copy_(step)
        torch.cuda.synchronize(dev)
    B, S = plan.B, net.image_size
    io = dict(anchors=torch.from_numpy(anchors_for_shape((S, S))).to(dev),
              gt_boxes=torch.zeros((B, kmax, 4), dtype=torch.float64, device=dev),
              gt_labels=torch.zeros((B, kmax), dtype=torch.int32, device=dev),
              gt_counts=torch.zeros((B,), dtype=torch.int32, device=dev),
              image_hw=torch.full((B, 2), float(S), dtype=torch.float64, device=dev),
              lr=torch.full((4,), float(tr.opt.lr), dtype=torch.float32, device=dev))
    regions = _Regions(dev)
    for k, t in io.items():
        regions.name(t, k)
    regions.name(plan.input_images.t, "images")
    regions.name(plan.loss_out.t, "losses")
    regions.name(net.flat, "weights")
    val_bases = {v.t.untyped_storage().data_ptr() for v in plan.vals if v.t is not None}

    calls = [("effdet_anchor_targets",
              (io["anchors"].data_ptr(), plan.N, io["gt_boxes"].data_ptr(), io["gt_labels"].data_ptr(),
               io["gt_counts"].data_ptr(), B, int(kmax), io["image_hw"].data_ptr(), net.num_classes, 0.4, 0.5,
               plan.reg_t.ptr, None, plan.state_t.ptr, plan.cls_t.ptr))]
    for op in plan.ops:
        if not hasattr(op.fn, "call"):
            raise ValueError("launch %s (%s) is not a plain C-ABI call" % (op.name, op.kind))
        calls.append(op.fn.call)
    start = 0 if getattr(tr, "train_backbone", False) else net.backbone_end
    n = net.flat.numel() - start
    # the learning rate lives in a one-float region the C side writes before every replay: the graph never changes
    calls.append(("effdet_sgd_momentum_step_dev_lr",
                  (net.flat.data_ptr() + 4 * start, net.grad_flat.data_ptr() + 4 * start,
                   net.velocity.data_ptr() + 4 * start, n, io["lr"].data_ptr(), float(tr.opt.momentum), 1.0)))
    # capture order: the plan's own multi-lane schedule; target assignment before and the optimizer after it are
    # ordered against everything (lane 0, the SGD waits for the tails of the other lanes)
    n_lanes = plan.graph_lanes()
    sched = plan.lane_schedule(0, len(plan.ops), n_lanes) if n_lanes > 1 else [(0, [])] * len(plan.ops)
    sched = [(0, [])] + [(lane, [j + 1 for j in waits]) for lane, waits in sched]
    tails = {}
    for i, (lane, _) in enumerate(sched):
        tails[lane] = i
    sched.append((0, sorted(i for lane, i in tails.items() if lane != 0)))
    ops = [struct.pack("<BH", lane, len(waits)) + b"".join(struct.pack("<I", j) for j in waits) +
           _encode_call(name, args, regions) for (name, args), (lane, waits) in zip(calls, sched)]

    torch.cuda.synchronize(dev)
    with open(path, "wb") as f:
        f.write(b"EFDPLAN1")
        f.write(struct.pack("<I", len(regions.order)))
        for r, (base, nbytes) in enumerate(regions.order):
            name = regions.names.get(r, "").encode()
            # state (anything that is not a planned activation buffer) ships with its contents; activations start
            # at zero (the padding channels of the per-level gradient buffers rely on it)
            has = 0 if (base in val_bases and r not in regions.names) else 1
            f.write(struct.pack("<QBH", nbytes, has, len(name)) + name)
            if has:
                view = torch.empty(0, dtype=torch.uint8, device=dev).set_(regions.storages[base], 0, (nbytes,))
                f.write(view.cpu().numpy().tobytes())
        f.write(struct.pack("<I", len(ops)))
        for o in ops:
            f.write(o)
    info = dict(regions=len(regions.order), ops=len(ops), lr=float(tr.opt.lr), decay=float(tr.opt.decay),
                momentum=float(tr.opt.momentum), batch=B, image_size=S, kmax=int(kmax), num_anchors=int(plan.N),
                u8_input=bool(u8_input), trainable_from=int(start),
                weights={k: [int(net.offsets[k]), list(net.weights[k].shape)] for k in net.offsets})
    import json
    with open(str(path) + ".json", "w") as f:       # sidecar: where each Keras weight lives in the "weights" region
        json.dump(info, f)
    return info
