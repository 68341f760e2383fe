"""A minimal `tensorflow` / `tensorflow.keras` stand-in backed by torch-CPU float64 ops, just large
enough for the reference's OWN graph-construction code (/root/reference/model.py, efficientnet.py,
tfkeras.py, layers.py, initializers.py -- imported UNMODIFIED) to build and run EfficientDet.

TEST INFRASTRUCTURE (used by tests/golden/make_golden_graph.py in the build container only).

What this pins: everything the reference's Python decides -- topology, layer names, channel /
SE widths, kernel sizes, strides, skip / drop conditions, which tensors feed which BiFPN node and
in which weight-index order, head depth / sharing / reshape / concat order, BN epsilons, the
activation at every site.  What it does NOT pin: the arithmetic inside TensorFlow's kernels
(SAME padding, BN formula, nearest upsampling ...): those are restated below from TF's documented
behaviour, once, independently of oracle/graph.py.

Functional-API subset: symbolic tensors (`KT`) record (layer, inputs); `Model.predict` evaluates
the recorded graph; `Model.__call__` re-applies a model to new symbolic tensors (shared heads).
Weights come from `WEIGHTS(key, shape)` with key = "[<nested model>/]<layer>/<weight>".
"""
import re
import sys
import types

import numpy as np
import torch
import torch.nn.functional as Fn

DT = torch.float64
WEIGHTS = None            # callable (key, shape) -> ndarray ; set by the user of this module
CREATED = []              # every layer object in creation order
INITIALIZERS = {}         # weight key -> initializer object (for known-answer checks)
_counters = {}


def _snake(name):
    s = re.sub(r"(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    s = re.sub(r"([a-z])([A-Z])", r"\1_\2", s).lower()
    return s


def _auto_name(cls):
    base = _snake(cls)
    n = _counters.get(base, 0)
    _counters[base] = n + 1
    return base if n == 0 else "%s_%d" % (base, n)


def reset():
    CREATED.clear()
    INITIALIZERS.clear()
    _counters.clear()


class KT:
    """Symbolic tensor."""

    def __init__(self, layer, inputs):
        self.layer, self.inputs = layer, inputs


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


class Layer:
    def __init__(self, name=None, trainable=True, dtype=None, **kwargs):
        self.name = name or _auto_name(type(self).__name__)
        self.trainable = trainable
        self.built = False
        self.scope = ""
        self.weights = {}
        CREATED.append(self)

    def add_weight(self, name=None, shape=None, initializer=None, trainable=True, dtype=None, **kw):
        key = self.scope + self.name + "/" + name
        shape = tuple(int(s) for s in shape)
        t = torch.from_numpy(np.asarray(WEIGHTS(key, shape), np.float64).reshape(shape))
        self.weights[key] = t
        INITIALIZERS[key] = initializer
        return t

    def build(self, input_shape):
        pass

    def __call__(self, inputs, **kwargs):
        return KT(self, inputs)

    def run(self, values):
        """values: torch tensor or list of them (mirrors the structure given to __call__)."""
        if not self.built:
            shp = [tuple(v.shape) for v in values] if isinstance(values, list) else tuple(values.shape)
            self.build(shp)
            self.built = True
        return self.call(values)

    def get_config(self):
        return {"name": self.name}


class InputLayer(Layer):
    pass


def Input(shape=None, tensor=None, **kw):
    return KT(InputLayer(name=kw.get("name")), None)


def _same_pad(x, k, s):
    """TF padding='same' on NCHW input (Appendix A.1)."""
    H, W = x.shape[-2:]

    def p(n):
        o = -(-n // s)
        t = max((o - 1) * s + k - n, 0)
        return t // 2, t - t // 2
    (t, b), (l, r) = p(H), p(W)
    return Fn.pad(x, (l, r, t, b))


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _activation(name):
    if name is None or name == "linear":
        return lambda x: x
    if callable(name):
        return name
    custom = utils.get_custom_objects()
    if name in custom:
        return custom[name]              # the reference's own swish (efficientnet.py:147-165)
    return {"relu": torch.relu, "sigmoid": torch.sigmoid}[name]


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None, use_bias=True,
                 kernel_initializer="glorot_uniform", bias_initializer="zeros", **kw):
        super().__init__(**kw)
        self.filters, self.k, self.s = int(filters), _pair(kernel_size), _pair(strides)
        assert self.k[0] == self.k[1] and self.s[0] == self.s[1] and padding == "same"
        self.activation, self.use_bias = _activation(activation), use_bias
        self.ki, self.bi = kernel_initializer, bias_initializer

    def build(self, shp):
        self.kernel = self.add_weight("kernel", (self.k[0], self.k[1], shp[-1], self.filters), self.ki)
        self.bias = self.add_weight("bias", (self.filters,), self.bi) if self.use_bias else None

    def call(self, x):
        y = Fn.conv2d(_same_pad(x.permute(0, 3, 1, 2), self.k[0], self.s[0]), self.kernel.permute(3, 2, 0, 1),
                      self.bias, stride=self.s[0])
        return self.activation(y.permute(0, 2, 3, 1))


class DepthwiseConv2D(Layer):
    def __init__(self, kernel_size, strides=1, padding="valid", use_bias=True,
                 depthwise_initializer="glorot_uniform", **kw):
        super().__init__(**kw)
        self.k, self.s = _pair(kernel_size), _pair(strides)
        assert self.k[0] == self.k[1] and self.s[0] == self.s[1] and padding == "same" and not use_bias
        self.ki = depthwise_initializer

    def build(self, shp):
        self.kernel = self.add_weight("depthwise_kernel", (self.k[0], self.k[1], shp[-1], 1), self.ki)

    def call(self, x):
        C = x.shape[-1]
        y = Fn.conv2d(_same_pad(x.permute(0, 3, 1, 2), self.k[0], self.s[0]), self.kernel.permute(2, 3, 0, 1),
                      None, stride=self.s[0], groups=C)
        return y.permute(0, 2, 3, 1)


class BatchNormalization(Layer):
    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, **kw):
        super().__init__(**kw)
        assert axis in (-1, 3)
        self.momentum, self.epsilon = momentum, epsilon

    def build(self, shp):
        C = shp[-1]
        self.gamma, self.beta = self.add_weight("gamma", (C,), "ones"), self.add_weight("beta", (C,), "zeros")
        self.mean = self.add_weight("moving_mean", (C,), "zeros")
        self.var = self.add_weight("moving_variance", (C,), "ones")

    def call(self, x):                       # inference phase (predict)
        return (x - self.mean) / torch.sqrt(self.var + self.epsilon) * self.gamma + self.beta


class Activation(Layer):
    def __init__(self, activation, **kw):
        super().__init__(**kw)
        self.fn = _activation(activation)

    def call(self, x):
        return self.fn(x)


class ReLU(Layer):
    def call(self, x):
        return torch.relu(x)


class UpSampling2D(Layer):
    def __init__(self, size=(2, 2), **kw):
        super().__init__(**kw)
        self.size = _pair(size)

    def call(self, x):                       # nearest
        return x.repeat_interleave(self.size[0], dim=1).repeat_interleave(self.size[1], dim=2)


class MaxPooling2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", **kw):
        super().__init__(**kw)
        self.pool, self.strides = _pair(pool_size), _pair(strides or pool_size)
        assert padding == "valid"

    def call(self, x):
        return Fn.max_pool2d(x.permute(0, 3, 1, 2), self.pool, self.strides).permute(0, 2, 3, 1)


class GlobalAveragePooling2D(Layer):
    def call(self, x):
        return x.mean(dim=(1, 2))


class Reshape(Layer):
    def __init__(self, target_shape, **kw):
        super().__init__(**kw)
        self.target = tuple(target_shape)

    def call(self, x):
        return x.reshape((x.shape[0],) + self.target)


class Dropout(Layer):
    def __init__(self, rate, noise_shape=None, seed=None, **kw):
        super().__init__(**kw)
        self.rate, self.noise_shape = rate, noise_shape

    def call(self, x):                       # inference phase: identity
        return x


class Add(Layer):
    def call(self, xs):
        y = xs[0]
        for t in xs[1:]:
            y = y + t
        return y


class Multiply(Layer):
    def call(self, xs):
        y = xs[0]
        for t in xs[1:]:
            y = y * t
        return y


class Concatenate(Layer):
    def __init__(self, axis=-1, **kw):
        super().__init__(**kw)
        self.axis = axis

    def call(self, xs):
        return torch.cat(xs, dim=self.axis)


class Lambda(Layer):
    def __init__(self, function, **kw):
        super().__init__(**kw)
        self.fn = function

    def call(self, x):
        return self.fn(x)


def add(inputs, **kw):
    return Add(**kw)(inputs)


def multiply(inputs, **kw):
    return Multiply(**kw)(inputs)


class Model(Layer):
    """Functional model: evaluates the recorded graph; callable on new symbolic tensors."""

    def __init__(self, inputs=None, outputs=None, name=None, **kw):
        super().__init__(name=name)
        self.inputs, self.outputs = _as_list(inputs), outputs
        self.inner = []
        seen = set()

        def walk(t):
            if id(t) in seen:
                return
            seen.add(id(t))
            if t.inputs is not None:
                for i in _as_list(t.inputs):
                    walk(i)
            if t.layer not in self.inner:
                self.inner.append(t.layer)
        for o in _as_list(outputs):
            walk(o)

    @property
    def layers(self):
        return list(self.inner)

    def __call__(self, inputs, **kwargs):
        for l in self.inner:                 # used as a layer of an outer model: nested weight scope
            l.scope = self.name + "/"
        return KT(self, inputs)

    def run(self, values):
        return self._evaluate(_as_list(values), None)

    def _evaluate(self, values, record):
        memo = {id(t): v for t, v in zip(self.inputs, values)}

        def ev(t):
            if id(t) in memo:
                return memo[id(t)]
            if isinstance(t.inputs, (list, tuple)):
                arg = [ev(i) for i in t.inputs]
            else:
                arg = ev(t.inputs)
            v = t.layer.run(arg)
            memo[id(t)] = v
            if record is not None:
                record[id(t)] = v
            return v
        outs = [ev(o) for o in _as_list(self.outputs)]
        return outs if isinstance(self.outputs, (list, tuple)) else outs[0]

    def predict(self, x, record=None):
        sys.setrecursionlimit(max(sys.getrecursionlimit(), 50000))
        vals = [torch.as_tensor(np.asarray(v)).to(DT) for v in _as_list(x)]
        with torch.no_grad():
            return self._evaluate(vals, record)

    predict_on_batch = predict

    def all_weights(self):
        d = {}
        for l in self.inner:
            if isinstance(l, Model):
                d.update(l.all_weights())
            d.update(l.weights)
        return d


# --------------------------------------------------------------------------- module assembly
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


class _Initializer:
    def __call__(self, shape, dtype=None):
        raise NotImplementedError

    def get_config(self):
        return {}


class _Constant(_Initializer):
    def __init__(self, value=0):
        self.value = value

    def __call__(self, shape, dtype=None):
        return torch.full(tuple(shape), float(self.value), dtype=DT)


class _RandomNormal(_Initializer):
    def __init__(self, mean=0.0, stddev=1.0, seed=None, **kw):
        self.mean, self.stddev = mean, stddev


_custom_objects = {}
utils = _module("tensorflow.keras.utils", get_custom_objects=lambda: _custom_objects)
backend = _module(
    "tensorflow.keras.backend", floatx=lambda: "float32", image_data_format=lambda: "channels_last",
    backend=lambda: "tensorflow", sigmoid=torch.sigmoid, is_keras_tensor=lambda t: isinstance(t, KT),
    shape=lambda t: t.shape)
layers = _module(
    "tensorflow.keras.layers", Layer=Layer, Input=Input, Conv2D=Conv2D, DepthwiseConv2D=DepthwiseConv2D,
    BatchNormalization=BatchNormalization, Activation=Activation, ReLU=ReLU, UpSampling2D=UpSampling2D,
    MaxPooling2D=MaxPooling2D, GlobalAveragePooling2D=GlobalAveragePooling2D, Reshape=Reshape,
    Dropout=Dropout, Add=Add, Multiply=Multiply, Concatenate=Concatenate, Lambda=Lambda, add=add,
    multiply=multiply)
models = _module("tensorflow.keras.models", Model=Model)
initializers = _module("tensorflow.keras.initializers", Initializer=_Initializer, constant=_Constant,
                       Constant=_Constant)
activations = _module("tensorflow.keras.activations", relu=torch.relu, sigmoid=torch.sigmoid)
keras = _module("tensorflow.keras", layers=layers, models=models, backend=backend, utils=utils,
                initializers=initializers, activations=activations, Model=Model)


def _reduce_sum(input_tensor=None, axis=None, **kw):
    t = torch.stack(list(input_tensor)) if isinstance(input_tensor, (list, tuple)) else input_tensor
    return t.sum() if axis is None else t.sum(dim=axis)


def _tf_constant(value, shape=None, name=None, dtype=None):
    return torch.full(tuple(shape), float(value), dtype=DT)


class _NameScope:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _where(cond, x=None, y=None, name=None):
    return torch.where(cond, x, y)


def _cast(x, dtype):
    if dtype in ("float32", torch.float32, torch.float64):
        return x.to(DT)
    if dtype in ("int32", torch.int32):
        return x.to(torch.int32)
    return x.to(dtype)


def _count_nonzero(x, dtype=None, **kw):
    return (x != 0).sum().to(DT)


def _maximum(a, b):
    return torch.maximum(torch.as_tensor(a, dtype=DT), torch.as_tensor(b, dtype=DT))


def _binary_crossentropy(target, output, from_logits=False):
    """keras.backend.binary_crossentropy of TF 1.15 / 2.0: clip to [eps, 1-eps] with
    eps = 1e-7, back to logits, sigmoid cross-entropy with logits."""
    assert not from_logits
    eps = 1e-7
    p = torch.clamp(output, eps, 1 - eps)
    z = torch.log(p / (1 - p))
    return torch.clamp(z, min=0) - z * target + torch.log1p(torch.exp(-torch.abs(z)))


backend.binary_crossentropy = _binary_crossentropy
tfmath = _module("tensorflow.math", reduce_sum=_reduce_sum, count_nonzero=_count_nonzero)
nn = _module("tensorflow.nn", swish=lambda x: x * torch.sigmoid(x), sigmoid=torch.sigmoid, relu=torch.relu)
compat_v1 = _module("tensorflow.compat.v1", random_normal_initializer=_RandomNormal, where=_where,
                    name_scope=_NameScope)
compat = _module("tensorflow.compat", v1=compat_v1)
tf = _module("tensorflow", keras=keras, nn=nn, compat=compat, reduce_sum=_reduce_sum, constant=_tf_constant,
             float32="float32", int32="int32", math=tfmath, abs=torch.abs, greater=torch.gt, equal=torch.eq,
             not_equal=torch.ne, cast=_cast, maximum=_maximum, zeros_like=torch.zeros_like)
backend.tf = tf

tf_utils = _module("tensorflow.python.keras.utils.tf_utils")
py_keras_utils = _module("tensorflow.python.keras.utils", tf_utils=tf_utils)
py_keras_backend = _module("tensorflow.python.keras.backend", is_keras_tensor=lambda t: isinstance(t, KT))
py_keras = _module("tensorflow.python.keras", utils=py_keras_utils, backend=py_keras_backend)
py = _module("tensorflow.python", keras=py_keras)
tf.python = py

imagenet_utils = _module("keras_applications.imagenet_utils",
                         _obtain_input_shape=lambda input_shape, **kw: input_shape,
                         decode_predictions=None, preprocess_input=None)
keras_applications = _module("keras_applications", imagenet_utils=imagenet_utils)


def install():
    sys.modules.update({
        "tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.layers": layers,
        "tensorflow.keras.models": models, "tensorflow.keras.backend": backend,
        "tensorflow.keras.utils": utils, "tensorflow.keras.initializers": initializers,
        "tensorflow.keras.activations": activations, "tensorflow.nn": nn, "tensorflow.math": tfmath, "tensorflow.compat": compat,
        "tensorflow.compat.v1": compat_v1, "tensorflow.python": py, "tensorflow.python.keras": py_keras,
        "tensorflow.python.keras.utils": py_keras_utils,
        "tensorflow.python.keras.utils.tf_utils": tf_utils,
        "tensorflow.python.keras.backend": py_keras_backend,
        "keras_applications": keras_applications, "keras_applications.imagenet_utils": imagenet_utils,
    })
