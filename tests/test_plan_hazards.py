"""Host logic of the launch plans, checked on the CPU on the REAL launch lists (no kernels run: the plan is built
with the library calls stubbed, which leaves the op list, the declared inputs / outputs and the buffer assignment
exactly as on the GPU; host addresses stand in for device addresses):

* liveness-based buffer reuse (engine.Plan._assign_buffers): two values never share a buffer while both are live;
* multi-lane CUDA-graph capture order (engine.Plan.dependencies / lane_schedule, train.TrainPlan.lane_hint): every
  read-after-write, write-after-write and write-after-read hazard between two launches on a (recycled) buffer is
  ordered by the lane order plus the cross-lane waits -- for the inference plan on 4 lanes, the training plans
  (frozen / trained backbone, weighted BiFPN, stochastic depth) on 2..6 lanes, and for the per-bucket segments the
  data-parallel step captures (a segment may rely on earlier segments being complete, never on later ones);
* memory OUTSIDE the planned values (parameter gradients that several launches accumulate into, shared scratch,
  stochastic-depth scales): found by scanning the bound arguments of every launch for addresses -- any such address
  two launches share must be ordered as well (engine.Plan.dependencies states this as an assumption; here it is
  checked).
"""
import bisect
import ctypes
import pytest
import torch


@pytest.fixture(scope="module")
def stubbed():
    """Library calls stubbed for the module: plans are built (values, launches, buffers, bound arguments) but nothing
    is launched."""
    from efficientdet_b200 import _lib, engine
    saved = (_lib.stream_ptr, _lib.call, engine.Plan.__init__)

    def structure_only(self, net, batch, reuse_buffers=True, keep_taps=False, u8_input=False):
        self.net, self.u8_input, self.B, self.dev, self.dtype = net, bool(u8_input), int(batch), net.device, net.dtype
        self.ops, self.vals, self.taps, self.keep_taps = [], [], {}, keep_taps
        self.reuse = reuse_buffers and not keep_taps
        self._keepalive, self.graph = [], None
        self._build()
        self._assign_buffers()
        for op in self.ops:             # binds entry point + arguments (op.fn.call); nothing is launched
            op.fn = op.make()

    _lib.stream_ptr = lambda device=None: 0
    _lib.call = lambda *a, **k: 0
    engine.Plan.__init__ = structure_only
    try:
        yield True
    finally:
        _lib.stream_ptr, _lib.call, engine.Plan.__init__ = saved


@pytest.fixture(scope="module")
def cpu_plans(stubbed):
    from efficientdet_b200 import engine, train
    from efficientdet_b200.model import efficientdet
    plans = {}
    m0 = efficientdet(0, num_classes=4, image_size=128, drop_connect_rate=0, just_training_model=True,
                      device="cpu", dtype="bf16")
    plans["infer_d0_b1"] = engine.Plan(m0.net, 1)
    plans["infer_d0_b1_u8"] = engine.Plan(m0.net, 1, u8_input=True)
    plans["train_d0_frozen"] = train.TrainPlan(m0.net, 2, train_backbone=False)
    m1 = efficientdet(1, num_classes=4, image_size=128, weighted_bifpn=True, just_training_model=True,
                      device="cpu", dtype="bf16")       # default drop_connect_rate: stochastic depth ops
    plans["train_d1_full"] = train.TrainPlan(m1.net, 2, train_backbone=True)
    m32 = efficientdet(0, num_classes=4, image_size=128, drop_connect_rate=0, just_training_model=True,
                       device="cpu", dtype="fp32")
    plans["train_d0_fp32_full"] = train.TrainPlan(m32.net, 2, train_backbone=True)
    return plans


def _buffer(v):
    return v.t.data_ptr()


def test_buffer_reuse_never_overlaps_live_ranges(cpu_plans):
    for name, p in cpu_plans.items():
        first = {}
        for idx, op in enumerate(p.ops):
            for v in op.inputs + op.outputs:
                first.setdefault(id(v), idx)
        by_buf = {}
        for v in p.vals:
            assert v.t is not None and v.t.numel() >= v.nbytes, name
            by_buf.setdefault(_buffer(v), []).append(v)
        shared = 0
        for vs in by_buf.values():
            live = sorted((first.get(id(v), -1), v.last_use, v) for v in vs)
            for (f0, l0, v0), (f1, l1, v1) in zip(live, live[1:]):
                shared += 1
                assert not v0.keep and not v1.keep, "%s: a kept value was recycled" % name
                assert l0 < f1, "%s: %s (live %d..%d) and %s (live %d..%d) share a buffer" % (
                    name, v0.name, f0, l0, v1.name, f1, l1)
        assert shared > 0, "%s: no buffer was reused (the test would be vacuous)" % name


def _happens_before(sched, begin):
    """bitset per launch of the launches guaranteed complete before it starts: its lane predecessor and its
    cross-lane waits, transitively (what stream order + event waits give inside a capture)."""
    hb, lane_last = [], {}
    for k, (lane, waits) in enumerate(sched):
        m = 0
        prev = lane_last.get(lane)
        if prev is not None:
            m |= hb[prev] | (1 << prev)
        for j in waits:
            assert begin <= j < begin + k, "wait on a launch that has not been enqueued"
            m |= hb[j - begin] | (1 << (j - begin))
        hb.append(m)
        lane_last[lane] = k
    return hb


def _check_segment(name, p, begin, end, n_lanes):
    sched = p.lane_schedule(begin, end, n_lanes)
    assert len(sched) == end - begin
    hb = _happens_before(sched, begin)
    last_w, readers, n_hazards = {}, {}, 0
    for i in range(begin, end):
        op = p.ops[i]
        rb = {_buffer(v) for v in op.inputs}
        wb = {_buffer(v) for v in op.outputs}
        need = set()
        for b in rb | wb:
            if b in last_w:
                need.add(last_w[b])                     # RAW / WAW
        for b in wb:
            need.update(readers.get(b, ()))             # WAR
        need.discard(i)
        for j in need:
            n_hazards += 1
            assert hb[i - begin] >> (j - begin) & 1, "%s (%d lanes): launch %d (%s %s) is not ordered after %d (%s %s)" % (
                name, n_lanes, i, op.kind, op.name, j, p.ops[j].kind, p.ops[j].name)
        for b in rb - wb:
            readers.setdefault(b, set()).add(i)
        for b in wb:
            last_w[b] = i
            readers[b] = set()
        lane, _ = sched[i - begin]
        hint = p.lane_hint(op, n_lanes)
        if isinstance(hint, int):
            assert lane == hint
        elif hint is not None:
            assert lane in hint
        assert 0 <= lane < n_lanes
    return n_hazards, len({l for l, _ in sched})


@pytest.mark.parametrize("n_lanes", [1, 2, 3, 4, 6])
def test_lane_schedule_orders_every_buffer_hazard(cpu_plans, n_lanes):
    for name, p in cpu_plans.items():
        n, used = _check_segment(name, p, 0, len(p.ops), n_lanes)
        assert n > len(p.ops) // 2
        if n_lanes >= 2 and name.startswith("train"):
            assert used >= 2, "%s: the weight gradients should have taken a second lane" % name


def test_lane_schedule_of_bucket_segments(cpu_plans):
    """Data-parallel capture: one graph per gradient bucket (engine.Plan.capture(bounds)); dependencies that reach
    into an earlier segment are dropped from the waits (the segments replay in order on one stream)."""
    from efficientdet_b200 import parallel
    for name in ("train_d0_frozen", "train_d1_full"):
        p = cpu_plans[name]
        buckets = parallel.plan_buckets(p.bucket_marks, p.net.flat.numel(), 1 << 12)
        bounds = [b[0] for b in buckets]
        assert bounds == sorted(bounds) and 0 < bounds[0] and bounds[-1] <= len(p.ops)
        assert len(buckets) >= 2, name
        # buckets partition the trainable range, in backward-production order (descending offsets)
        his = [b[2] for b in buckets]
        los = [b[1] for b in buckets]
        assert his[0] == p.net.flat.numel() and all(los[k] == his[k + 1] for k in range(len(buckets) - 1))
        begin = 0
        for end in bounds[:-1] + [len(p.ops)]:
            _check_segment(name, p, begin, end, 4)
            begin = end


def test_barrier_launches_are_ordered_against_everything(cpu_plans):
    """A launch that declares neither inputs nor outputs (the stochastic-depth mask draw) must follow every earlier
    launch and precede every later one."""
    p = cpu_plans["train_d1_full"]
    barriers = [i for i, op in enumerate(p.ops) if not op.inputs and not op.outputs]
    assert barriers, "the D1 training plan should draw stochastic-depth masks"
    sched = p.lane_schedule(0, len(p.ops), 4)
    hb = _happens_before(sched, 0)
    for b in barriers:
        assert hb[b] == (1 << b) - 1
        for i in range(b + 1, len(p.ops)):
            assert hb[i] >> b & 1


def _addresses(a, out):
    if isinstance(a, bool):
        return
    if isinstance(a, int):
        if a > (1 << 24):
            out.add(a)
    elif isinstance(a, (list, tuple, ctypes.Array)):
        for x in a:
            _addresses(x, out)
    elif isinstance(a, ctypes.Structure):
        for f in a._fields_:
            _addresses(getattr(a, f[0]), out)
    elif isinstance(a, ctypes.c_void_p):
        if a.value:
            out.add(a.value)
    elif hasattr(a, "_obj"):                # ctypes.byref(...)
        _addresses(a._obj, out)


@pytest.mark.parametrize("n_lanes", [2, 4, 6])
def test_untracked_memory_shared_by_launches_is_ordered(cpu_plans, n_lanes):
    for name, p in cpu_plans.items():
        size = {}
        for v in p.vals:
            size[v.t.data_ptr()] = max(size.get(v.t.data_ptr(), 0), v.t.numel())
        starts = sorted(size)

        def planned(x):
            k = bisect.bisect_right(starts, x) - 1
            return k >= 0 and x < starts[k] + size[starts[k]]
        flat = p.net.flat
        w_lo, w_hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size()
        sched = p.lane_schedule(0, len(p.ops), n_lanes)
        hb = _happens_before(sched, 0)
        touched, n_bound = {}, 0
        for i, op in enumerate(p.ops):
            call = getattr(op.fn, "call", None)
            assert call is not None, "%s: launch %d (%s) does not expose its bound arguments" % (name, i, op.kind)
            found = set()
            _addresses(call[1], found)
            n_bound += len(found)
            for x in found:
                # weights are read-only inside a step (the BatchNorm moving statistics have one writer and no
                # reader in training mode); planned values are covered by the hazard test above
                if not planned(x) and not (w_lo <= x < w_hi):
                    touched.setdefault(x, []).append(i)
        assert n_bound > 2 * len(p.ops)
        shared = [ops for ops in touched.values() if len(ops) > 1]
        assert shared or not name.startswith("train"), name
        for ops in shared:
            for a, i in enumerate(ops):
                for j in ops[a + 1:]:
                    assert hb[j] >> i & 1, (
                        "%s (%d lanes): launches %d (%s %s, lane %d) and %d (%s %s, lane %d) share memory outside the "
                        "planned values but are not ordered -- declare it as a value, or pin both to one lane "
                        "(if it is read-only for both, exempt it here)" % (
                            name, n_lanes, i, p.ops[i].kind, p.ops[i].name, sched[i][0],
                            j, p.ops[j].kind, p.ops[j].name, sched[j][0]))


@pytest.mark.parametrize("phi", range(7))
def test_every_model_size_lowers_and_schedules(stubbed, phi):
    """D0..D6, weighted and plain BiFPN, bf16 and fp32, inference (float / uint8 input) and training (frozen / trained
    backbone): the lowering runs to the end (every descriptor is built, every argument bound) and the 4-lane order
    covers every hazard.  The GPU tests run a subset of these combinations; this is all of them."""
    from efficientdet_b200 import engine, plan_export, train
    from efficientdet_b200.model import efficientdet

    class _AnyRegion:
        def locate(self, ptr):
            return 0, ptr
    for weighted in (False, True):
        for dtype in ("bf16", "fp32"):
            m = efficientdet(phi, num_classes=3 + phi, image_size=128, weighted_bifpn=weighted,
                             just_training_model=True, device="cpu", dtype=dtype)
            for B in (1, 3):
                p = engine.Plan(m.net, B, u8_input=(B == 3))
                _check_segment("infer D%d" % phi, p, 0, len(p.ops), 4)
            n_ops = []
            for full in (False, True):
                p = train.TrainPlan(m.net, 2, train_backbone=full, u8_input=full)
                _check_segment("train D%d" % phi, p, 0, len(p.ops), 4)
                n_ops.append(len(p.ops))
                # every launch of a training plan must serialise into a compiled plan (plan_export / csrc/replay.cu):
                # a plain C-ABI call whose recorded arguments match the signature table the replay thunks are
                # generated from
                for op in p.ops:
                    name, args = op.fn.call
                    blob = plan_export._encode_call(name, args, _AnyRegion())
                    assert blob[2:2 + len(name)] == name.encode()
            assert n_ops[1] > n_ops[0]          # the trained backbone adds its backward launches
