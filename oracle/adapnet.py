"""Oracle: the Adapnet expert (TEST INFRASTRUCTURE ONLY, see oracle/__init__).

Follows xview/models/adapnet.py:12-173 (block_a, block_b, adapnet) on top of
xview/models/custom_layers.py:71-139 (conv2d / deconv2d with batch normalisation between the
convolution and the activation).  Test-time graph only (`is_training=False`: batch norm uses its
moving statistics), which is what the fusion models build (basic_fusion_model.py:13-16).
Activations NHWC numpy float32, convolutions in torch-CPU fp32.

PARITY UNPINNED at the TensorFlow boundary, like oracle/fcn.py: TF 'SAME' padding of strided and
dilated convolutions is restated from its documented rule
(out = ceil(in / s), pad_total = max((out - 1) s + (k - 1) d + 1 - in, 0), the smaller half first).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .fcn import _batchnorm, _t, bilinear_filter, deconv2d, max_pool2x2

# (name, kind, args) in execution order, adapnet.py:124-153
# block_a args: (intermediate, filters, strides, shortcut_conv)
# block_b args: (filters_1, filters_2, filters_3, dilation1, dilation2, shortcut_conv)
BLOCKS = [
    ('block_layer_1', 'a', (64, 256, 1, True)),
    ('block_layer_2', 'a', (64, 256, 1, False)),
    ('block_layer_3', 'a', (64, 256, 1, False)),
    ('block_layer_4', 'a', (128, 512, 2, True)),
    ('block_layer_5', 'a', (128, 512, 1, False)),
    ('block_layer_6', 'a', (128, 512, 1, False)),
    ('block_layer_7', 'b', (128, 64, 512, 1, 2, False)),
    ('block_layer_8', 'a', (256, 1024, 2, True)),
    ('block_layer_9', 'a', (256, 1024, 1, False)),
    ('block_layer_10', 'b', (256, 256, 1024, 1, 2, False)),
    ('block_layer_11', 'b', (256, 256, 1024, 1, 4, False)),
    ('block_layer_12', 'b', (256, 256, 1024, 1, 8, False)),
    ('block_layer_13', 'b', (256, 256, 1024, 1, 16, False)),
    ('block_layer_14', 'b', (512, 512, 2048, 2, 4, True)),
    ('block_layer_15', 'b', (512, 512, 2048, 2, 8, False)),
    ('block_layer_16', 'b', (512, 512, 2048, 2, 16, False)),
]


def adapnet_param_shapes(prefix, cin, num_units, num_classes):
    """Variable names + shapes: every conv has `<scope>/kernel`, the four batch-norm tensors
    under the same scope (custom_layers.py:114-116,132-134 pass name=<scope>) and a
    `<scope>/bias` unless use_bias=False (all block convolutions, adapnet.py:35-36,78-79; the
    transposed convs, custom_layers.py:80)."""
    shapes = {}

    def conv(scope, k, ci, co, bias):
        shapes[scope + '/kernel'] = (k, k, ci, co)
        if bias:
            shapes[scope + '/bias'] = (co,)
        for leaf in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
            shapes['%s/%s' % (scope, leaf)] = (co,)

    conv(prefix + '/block_0_1', 3, cin, 64, True)
    conv(prefix + '/block_0_2', 7, 64, 64, True)
    c = 64
    for name, kind, args in BLOCKS:
        scope = '%s/%s' % (prefix, name)
        if kind == 'a':
            mid, out, _, shortcut = args
            conv(scope + '/stage_1', 1, c, mid, False)
            conv(scope + '/stage_2', 3, mid, mid, False)
            conv(scope + '/stage_3', 1, mid, out, False)
        else:
            f1, f2, out, _, _, shortcut = args
            conv(scope + '/stage_1', 1, c, f1, False)
            conv(scope + '/stage_2_1', 3, f1, f2 // 2, False)
            conv(scope + '/stage_2_2', 3, f1, f2 // 2, False)
            conv(scope + '/stage_3', 1, f2, out, False)
        if shortcut:
            conv(scope + '/shortcut', 1, c, out, False)
        c = out
        if name == 'block_layer_7':
            conv(prefix + '/shortcut', 1, c, num_units, True)
    conv(prefix + '/first_deconvolution_conv', 1, 2048, 2048, True)
    # transposed convs: kernel [kh, kw, Cout, Cin] (custom_layers.py:92), no bias, batch norm
    for scope, k, co, ci in ((prefix + '/first_deconvolution_upconv', 4, num_units, 2048),
                             (prefix + '/second_deconvolution_upconv', 16, num_classes,
                              num_units)):
        shapes[scope + '/kernel'] = (k, k, co, ci)
        for leaf in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
            shapes['%s/%s' % (scope, leaf)] = (co,)
    return shapes


def adapnet_params(prefix, cin, num_units, num_classes, rng, gain=1.0, trained_like=True):
    """Random init of the architecture.  Kernels: Glorot uniform (tf.layers default) times
    `gain`; transposed convs: the bilinear kernel of custom_layers.py:8-25 (plus, when
    `trained_like`, a small dense perturbation so that off-diagonal taps are exercised); batch
    norm: identity statistics, or perturbed ones when `trained_like`."""
    params = {}
    for name, shape in adapnet_param_shapes(prefix, cin, num_units, num_classes).items():
        leaf = name.split('/')[-1]
        if leaf == 'kernel' and 'upconv' in name:
            w = bilinear_filter_rect(shape)
            if trained_like:
                w = w + 0.02 * rng.standard_normal(shape) / np.sqrt(shape[3])
            params[name] = w.astype(np.float32)
        elif leaf == 'kernel':
            fan_in = shape[0] * shape[1] * shape[2]
            fan_out = shape[0] * shape[1] * shape[3]
            limit = gain * np.sqrt(6.0 / (fan_in + fan_out))
            params[name] = rng.uniform(-limit, limit, size=shape).astype(np.float32)
        elif leaf == 'bias':
            params[name] = ((0.05 * rng.standard_normal(shape)) if trained_like
                            else np.zeros(shape)).astype(np.float32)
        elif leaf in ('gamma', 'moving_variance'):
            params[name] = ((1.0 + 0.2 * rng.random(shape)) if trained_like
                            else np.ones(shape)).astype(np.float32)
        else:
            params[name] = ((0.1 * rng.standard_normal(shape)) if trained_like
                            else np.zeros(shape)).astype(np.float32)
    return params


def bilinear_filter_rect(filter_shape):
    """custom_layers.py:8-25 for Cout != Cin: the reference loops `i` over filter_shape[2] and
    writes weights[:, :, i, i], which needs i < filter_shape[3] as well - with
    kernel_dims = [k, k, filters, Cin] (custom_layers.py:92) that holds for both Adapnet
    transposed convs (num_units <= 2048, num_classes <= num_units)."""
    if filter_shape[2] <= filter_shape[3]:
        return bilinear_filter(filter_shape)
    raise ValueError('bilinear initialiser needs Cout <= Cin')


def same_padding(size, k, stride, dilation):
    """TF 'SAME': (pad_before, pad_after) along one axis."""
    out = -(-size // stride)
    total = max((out - 1) * stride + (k - 1) * dilation + 1 - size, 0)
    return total // 2, total - total // 2


def conv_bn(x, params, scope, stride=1, dilation=1, activation=True):
    """custom_layers.py:124-139 with batch_normalization=True: conv 'same' (+bias if present)
    -> batch norm (moving statistics) -> ReLU."""
    w = params[scope + '/kernel'].astype(x.dtype)
    b = params.get(scope + '/bias')
    k = w.shape[0]
    ph = same_padding(x.shape[1], k, stride, dilation)
    pw = same_padding(x.shape[2], k, stride, dilation)
    xt = F.pad(_t(x).permute(0, 3, 1, 2), (pw[0], pw[1], ph[0], ph[1]))
    y = F.conv2d(xt, _t(w).permute(3, 2, 0, 1), None if b is None else _t(b.astype(x.dtype)),
                 stride=stride, dilation=dilation)
    y = y.permute(0, 2, 3, 1).contiguous().numpy()
    y = _batchnorm(y, params, scope).astype(x.dtype)
    return np.maximum(y, 0) if activation else y


def block_a(x, params, scope, mid, out, stride, shortcut_conv):
    """adapnet.py:12-49."""
    s1 = conv_bn(x, params, scope + '/stage_1', stride=stride)
    s2 = conv_bn(s1, params, scope + '/stage_2')
    s3 = conv_bn(s2, params, scope + '/stage_3')
    shortcut = conv_bn(x, params, scope + '/shortcut', stride=stride) if shortcut_conv else x
    return np.maximum(s3 + shortcut, 0)


def block_b(x, params, scope, f1, f2, out, dilation1, dilation2, shortcut_conv):
    """adapnet.py:52-96: stage 2 is two atrous 3x3 convs, concatenated on the channel axis."""
    s1 = conv_bn(x, params, scope + '/stage_1')
    s21 = conv_bn(s1, params, scope + '/stage_2_1', dilation=dilation1)
    s22 = conv_bn(s1, params, scope + '/stage_2_2', dilation=dilation2)
    s3 = conv_bn(np.concatenate([s21, s22], axis=3), params, scope + '/stage_3')
    shortcut = conv_bn(x, params, scope + '/shortcut') if shortcut_conv else x
    return np.maximum(s3 + shortcut, 0)


def adapnet(x, params, prefix, num_units, num_classes):
    """adapnet.py:99-173: dict of the block outputs, 'score' = [N,H,W,num_classes]."""
    layers = {}
    layers['block_0_1'] = conv_bn(x, params, prefix + '/block_0_1')
    layers['block_0_2'] = conv_bn(layers['block_0_1'], params, prefix + '/block_0_2', stride=2)
    layers['block_0_pool'] = max_pool2x2(layers['block_0_2'])
    y = layers['block_0_pool']
    for index, (name, kind, args) in enumerate(BLOCKS, start=1):
        scope = '%s/%s' % (prefix, name)
        y = block_a(y, params, scope, *args) if kind == 'a' else block_b(y, params, scope, *args)
        layers['block_%d' % index] = y
        if index == 7:
            layers['shortcut'] = conv_bn(y, params, prefix + '/shortcut', activation=False)
    d = conv_bn(y, params, prefix + '/first_deconvolution_conv')
    layers['deconv_1'] = deconv2d(d, params, prefix + '/first_deconvolution_upconv', 2,
                                  activation=False, batchnorm=True)
    layers['merge'] = layers['deconv_1'] + layers['shortcut']
    layers['score'] = deconv2d(layers['merge'], params, prefix + '/second_deconvolution_upconv',
                               8, activation=False, batchnorm=True)
    return layers
