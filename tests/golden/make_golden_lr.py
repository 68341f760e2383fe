"""Generates tests/golden/lr_schedule.json by EXECUTING THE REFERENCE's own utils/lr_schedule.py
(get_cosine_decay_with_linear_warmup, /root/reference/utils/lr_schedule.py:5-68, imported unmodified).  Its two
TensorFlow dependencies are stood in for here: tf.keras.experimental.CosineDecay (restated from its documented formula:
lr0 * ((1 - alpha) * 0.5 * (1 + cos(pi * min(step, decay_steps) / decay_steps)) + alpha)) and
tf.keras.callbacks.LearningRateScheduler (holds the schedule function).  What the fixture pins is the reference's
own wiring: epoch_number = epoch_index + 1, the warm-up slope, the switch epoch and its strict `>` comparison.
Build container only.  Run:  python tests/golden/make_golden_lr.py
"""
import json
import math
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


class CosineDecay:
    def __init__(self, initial_learning_rate, decay_steps, alpha=0.0, name=None):
        self.lr0, self.steps, self.alpha = initial_learning_rate, decay_steps, alpha

    def __call__(self, step):
        step = min(step, self.steps)
        return self.lr0 * ((1 - self.alpha) * 0.5 * (1 + math.cos(math.pi * step / self.steps)) + self.alpha)


class LearningRateScheduler:
    def __init__(self, schedule, verbose=0):
        self.schedule = schedule


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


backend = _module("tensorflow.keras.backend", get_value=lambda x: x)
keras = _module("tensorflow.keras", backend=backend, experimental=_module("tensorflow.keras.experimental", CosineDecay=CosineDecay),
                callbacks=_module("tensorflow.keras.callbacks", LearningRateScheduler=LearningRateScheduler))
tf = _module("tensorflow", keras=keras)
sys.modules.update({"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.backend": backend})
sys.path.insert(0, "/root/reference")
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_lr_schedule", "/root/reference/utils/lr_schedule.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

cases = [dict(total_epochs=100), dict(total_epochs=100, learning_rate_max=0.08, warmup_percent=0.05, alpha=0.001),
         dict(total_epochs=30, learning_rate_start=0.001, learning_rate_max=0.02, warmup_percent=0.1, alpha=0.01),
         dict(total_epochs=7, learning_rate_max=0.5, warmup_percent=0.3, alpha=0.0),
         dict(total_epochs=500, learning_rate_start=0.01, learning_rate_max=0.01, warmup_percent=0.02)]
out = []
for kw in cases:
    cb = ref.get_cosine_decay_with_linear_warmup(**kw)
    n = kw["total_epochs"]
    out.append(dict(kwargs=kw, lr=[float(cb.schedule(e, None)) for e in range(n + 3)]))
json.dump(out, open(os.path.join(HERE, "lr_schedule.json"), "w"))
print([(c["kwargs"], c["lr"][:3], c["lr"][-1]) for c in out])
