"""Minimal stand-in for the keras.layers.Layer protocol the reference's layers follow
(__init__(**kwargs) incl. name, build, call, __call__, compute_output_shape, get_config,
get_weights/set_weights, trainable)."""
import collections

_name_counts = collections.defaultdict(int)


def _auto_name(cls_name):
    # keras: CamelCase -> snake_case, then _1, _2 ... ("w_bi_fpn_add", "w_bi_fpn_add_1", ...)
    s = ""
    for i, ch in enumerate(cls_name):
        if ch.isupper() and i and (not cls_name[i - 1].isupper() or
                                   (i + 1 < len(cls_name) and cls_name[i + 1].islower())):
            s += "_"
        s += ch.lower()
    n = _name_counts[s]
    _name_counts[s] += 1
    return s if n == 0 else "%s_%d" % (s, n)


class Layer:
    def __init__(self, name=None, trainable=True, dtype=None, **kwargs):
        if kwargs:
            raise TypeError("Keyword argument not understood: %s" % sorted(kwargs)[0])
        self.name = name or _auto_name(type(self).__name__)
        self.trainable = trainable
        self.built = False
        self._weights = collections.OrderedDict()     # name -> torch tensor
        self._dtype = dtype

    def add_weight(self, name, shape, value, trainable=True):
        self._weights[name] = value
        return value

    def build(self, input_shape):
        self.built = True

    def call(self, inputs, **kwargs):
        raise NotImplementedError

    def __call__(self, inputs, **kwargs):
        if not self.built:
            shapes = [tuple(getattr(x, "shape", ())) for x in inputs] \
                if isinstance(inputs, (list, tuple)) else tuple(getattr(inputs, "shape", ()))
            self.build(shapes)
            self.built = True
        return self.call(inputs, **kwargs)

    def compute_output_shape(self, input_shape):
        return input_shape

    def get_config(self):
        return {"name": self.name, "trainable": self.trainable}

    def get_weights(self):
        return [w.detach().cpu().numpy() for w in self._weights.values()]

    def set_weights(self, weights):
        import torch
        if len(weights) != len(self._weights):
            raise ValueError("layer %s expects %d weights, got %d" %
                             (self.name, len(self._weights), len(weights)))
        for (k, old), new in zip(list(self._weights.items()), weights):
            t = torch.as_tensor(new, dtype=old.dtype, device=old.device).reshape(old.shape)
            old.copy_(t)
