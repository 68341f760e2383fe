"""EfficientNet backbone description -- mirror of the reference's efficientnet.py public surface
(BlockArgs, DEFAULT_BLOCKS_ARGS, round_filters, round_repeats, EfficientNet, EfficientNetB0..B7)
for the feature-extractor use the reference makes of it (efficientnet.py:309-470 returns the five
stride-2..32 features, never a classifier top).

There is no symbolic Keras graph here: `EfficientNet(...)` returns a `BackboneSpec` (stem +
MBConv block records with Keras layer names) that model.efficientdet() lowers to a static
launch plan of the library's CUDA kernels.
"""
import collections
import math
import string

BlockArgs = collections.namedtuple('BlockArgs', [
    'kernel_size', 'num_repeat', 'input_filters', 'output_filters',
    'expand_ratio', 'id_skip', 'strides', 'se_ratio'])
BlockArgs.__new__.__defaults__ = (None,) * len(BlockArgs._fields)

DEFAULT_BLOCKS_ARGS = [
    BlockArgs(3, 1, 32, 16, 1, True, [1, 1], 0.25),
    BlockArgs(3, 2, 16, 24, 6, True, [2, 2], 0.25),
    BlockArgs(5, 2, 24, 40, 6, True, [2, 2], 0.25),
    BlockArgs(3, 3, 40, 80, 6, True, [2, 2], 0.25),
    BlockArgs(5, 3, 80, 112, 6, True, [1, 1], 0.25),
    BlockArgs(5, 4, 112, 192, 6, True, [2, 2], 0.25),
    BlockArgs(3, 1, 192, 320, 6, True, [1, 1], 0.25),
]

BN_EPSILON = 1e-3      # keras BatchNormalization default, efficientnet.py:233-236 passes none
BN_MOMENTUM = 0.99


def round_filters(filters, width_coefficient, depth_divisor):
    """Round number of filters based on width multiplier (efficientnet.py:191-201)."""
    filters *= width_coefficient
    new_filters = max(depth_divisor,
                      int(filters + depth_divisor / 2) // depth_divisor * depth_divisor)
    if new_filters < 0.9 * filters:
        new_filters += depth_divisor
    return int(new_filters)


def round_repeats(repeats, depth_coefficient):
    """Round number of repeats based on depth multiplier (efficientnet.py:204-207)."""
    return int(math.ceil(depth_coefficient * repeats))


class MBConv(collections.namedtuple("MBConv", [
        "prefix", "kernel_size", "stride", "input_filters", "output_filters", "expand_ratio",
        "se_filters", "has_skip", "drop_rate"])):
    @property
    def mid_filters(self):
        return self.input_filters * self.expand_ratio


class BackboneSpec:
    def __init__(self, stem_filters, blocks, feature_after, freeze_bn, model_name):
        self.stem_filters = stem_filters
        self.blocks = blocks                  # [MBConv]
        self.feature_after = feature_after    # indices into blocks after which C1..C5 are tapped
        self.freeze_bn = freeze_bn
        self.model_name = model_name

    def keras_layer_names(self):
        """Keras layer names in creation order (used for model.layers / freeze ranges,
        train.py:337-339, train_tpu.py:272-274)."""
        names = ["stem_conv", "stem_bn", "stem_activation"]
        for b in self.blocks:
            p = b.prefix
            if b.expand_ratio != 1:
                names += [p + "expand_conv", p + "expand_bn", p + "expand_activation"]
            names += [p + "dwconv", p + "bn", p + "activation"]
            names += [p + "se_squeeze", p + "se_reshape", p + "se_reduce", p + "se_expand",
                      p + "se_excite"]
            names += [p + "project_conv", p + "project_bn"]
            if b.has_skip:
                if b.drop_rate and b.drop_rate > 0:
                    names.append(p + "drop")
                names.append(p + "add")
        return names


def EfficientNet(width_coefficient, depth_coefficient, default_resolution, dropout_rate=0.2,
                 drop_connect_rate=0.2, depth_divisor=8, blocks_args=DEFAULT_BLOCKS_ARGS,
                 model_name='efficientnet', include_top=True, weights='imagenet',
                 input_tensor=None, input_shape=None, pooling=None, classes=1000,
                 freeze_bn=False, **kwargs):
    """Builds the backbone description.  `weights`, `include_top`, `pooling`, `classes`,
    `input_tensor`, `input_shape` are accepted for signature parity; like the reference
    (efficientnet.py:309-470) no classifier top is built and no weights are downloaded."""
    for key in kwargs:
        if key not in ('backend', 'layers', 'models', 'utils'):
            raise TypeError('Invalid keyword argument: %s' % key)
    import os
    if not (weights in {'imagenet', None} or os.path.exists(weights)):
        raise ValueError('The `weights` argument should be either `None` (random initialization), '
                         '`imagenet` (pre-training on ImageNet), or the path to the weights file '
                         'to be loaded.')
    if weights == 'imagenet' and include_top and classes != 1000:
        raise ValueError('If using `weights` as `"imagenet"` with `include_top` as true, '
                         '`classes` should be 1000')
    num_blocks_total = sum(b.num_repeat for b in blocks_args)     # unscaled, as in the reference
    blocks, taps, block_num = [], [], 0
    for idx, ba in enumerate(blocks_args):
        assert ba.num_repeat > 0
        cin = round_filters(ba.input_filters, width_coefficient, depth_divisor)
        cout = round_filters(ba.output_filters, width_coefficient, depth_divisor)
        rep = round_repeats(ba.num_repeat, depth_coefficient)
        has_se = ba.se_ratio is not None and 0 < ba.se_ratio <= 1
        if not has_se:
            raise ValueError("blocks without squeeze-excite are not supported by this build")
        for r in range(rep):
            ci = cin if r == 0 else cout
            st = ba.strides[0] if r == 0 else 1
            blocks.append(MBConv(
                prefix='block{}{}_'.format(idx + 1, string.ascii_lowercase[r]),
                kernel_size=ba.kernel_size, stride=st, input_filters=ci, output_filters=cout,
                expand_ratio=ba.expand_ratio, se_filters=max(1, int(ci * ba.se_ratio)),
                has_skip=bool(ba.id_skip and st == 1 and ci == cout),
                drop_rate=drop_connect_rate * float(block_num) / num_blocks_total))
            block_num += 1
        if idx < len(blocks_args) - 1 and blocks_args[idx + 1].strides[0] == 2:
            taps.append(len(blocks) - 1)
        elif idx == len(blocks_args) - 1:
            taps.append(len(blocks) - 1)
    return BackboneSpec(round_filters(32, width_coefficient, depth_divisor), blocks, taps,
                        freeze_bn, model_name)


def _variant(w, d, res, drop, name):
    def f(include_top=True, weights='imagenet', input_tensor=None, input_shape=None, pooling=None,
          classes=1000, **kwargs):
        return EfficientNet(w, d, res, drop, model_name=name, include_top=include_top,
                            weights=weights, input_tensor=input_tensor, input_shape=input_shape,
                            pooling=pooling, classes=classes, **kwargs)
    f.__name__ = name
    return f


EfficientNetB0 = _variant(1.0, 1.0, 224, 0.2, 'efficientnet-b0')
EfficientNetB1 = _variant(1.0, 1.1, 240, 0.2, 'efficientnet-b1')
EfficientNetB2 = _variant(1.1, 1.2, 260, 0.3, 'efficientnet-b2')
EfficientNetB3 = _variant(1.2, 1.4, 300, 0.3, 'efficientnet-b3')
EfficientNetB4 = _variant(1.4, 1.8, 380, 0.4, 'efficientnet-b4')
EfficientNetB5 = _variant(1.6, 2.2, 456, 0.4, 'efficientnet-b5')
EfficientNetB6 = _variant(1.8, 2.6, 528, 0.5, 'efficientnet-b6')
EfficientNetB7 = _variant(2.0, 3.1, 600, 0.5, 'efficientnet-b7')
