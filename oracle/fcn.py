"""Oracle: the per-modality VGG16-FCN expert (TEST INFRASTRUCTURE ONLY, see oracle/__init__).

Follows xview/models/simple_fcn.py:10-170 (encoder / decoder / fcn),
xview/models/custom_layers.py:8-25,71-139 (bilinear init, deconv2d, conv2d) and
xview/models/basic_fusion_model.py:9-23 (test_pipeline).  Activations are NHWC numpy
float32 at the interface; convolutions run in torch-CPU fp32 (or float64 when the input
arrays are float64 - the "fp64 twin" used for tolerance studies).

PARITY UNPINNED at the TensorFlow boundary (no TF in the image); op semantics are the
documented tf.layers ones restated in SURVEY.md Appendix A.
"""
import numpy as np
import torch
import torch.nn.functional as F

CONV_LAYERS = [  # (name, Cout) in execution order, simple_fcn.py:39-67
    ('conv1_1', 64), ('conv1_2', 64),
    ('conv2_1', 128), ('conv2_2', 128),
    ('conv3_1', 256), ('conv3_2', 256), ('conv3_3', 256),
    ('conv4_1', 512), ('conv4_2', 512), ('conv4_3', 512),
    ('conv5_1', 512), ('conv5_2', 512), ('conv5_3', 512)]


def bilinear_kernel_1d(size):
    """custom_layers.py:13-21: f=ceil(k/2), c=(2f-1-f%2)/(2f), w[x]=1-|x/f-c|."""
    factor = np.ceil(size / 2.0)
    center = (2 * factor - 1 - factor % 2) / (2.0 * factor)
    return np.array([1 - abs(x / factor - center) for x in range(size)])


def bilinear_filter(filter_shape):
    """custom_layers.py:8-25: [kh,kw,Cout,Cin] kernel, bilinear on the channel diagonal."""
    # the reference derives factor/center from the width only and uses them on both axes
    factor = np.ceil(filter_shape[0] / 2.0)
    center = (2 * factor - 1 - factor % 2) / (2.0 * factor)
    k1 = np.array([1 - abs(x / factor - center) for x in range(filter_shape[0])])
    k2 = np.array([1 - abs(y / factor - center) for y in range(filter_shape[1])])
    bilinear = np.outer(k1, k2)
    weights = np.zeros(filter_shape)
    for i in range(filter_shape[2]):
        weights[:, :, i, i] = bilinear
    return weights.astype(np.float32)


def fcn_param_shapes(prefix, cin, num_units, num_classes, batchnorm=False):
    """Variable names + shapes of one expert; SURVEY.md Appendix B /
    `Synthia Rand Cityscapes Examples.ipynb`:898-931 (34 tensors without batchnorm)."""
    shapes = {}
    c = cin
    for name, cout in CONV_LAYERS:
        shapes['%s/%s/kernel' % (prefix, name)] = (3, 3, c, cout)
        shapes['%s/%s/bias' % (prefix, name)] = (cout,)
        c = cout
    for name in ('score_conv4', 'score_conv5'):
        shapes['%s/%s/kernel' % (prefix, name)] = (1, 1, 512, num_units)
        shapes['%s/%s/bias' % (prefix, name)] = (num_units,)
    shapes['%s/upscore_conv5/kernel' % prefix] = (4, 4, num_units, num_units)
    shapes['%s/upscore/kernel' % prefix] = (16, 16, num_units, num_units)
    shapes['%s/score/kernel' % prefix] = (1, 1, num_units, num_classes)
    shapes['%s/score/bias' % prefix] = (num_classes,)
    if batchnorm:
        for name, cout in CONV_LAYERS + [('score_conv4', num_units),
                                         ('score_conv5', num_units),
                                         ('upscore_conv5', num_units),
                                         ('upscore', num_units),
                                         ('score', num_classes)]:
            for v in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
                shapes['%s/%s/%s' % (prefix, name, v)] = (cout,)
    return shapes


def glorot_fcn_params(prefix, cin, num_units, num_classes, rng, gain=1.0, bias_scale=0.0,
                      batchnorm=False):
    """Random init of the named architecture: tf.layers default = Glorot-uniform kernels
    and zero biases (custom_layers.py:131,138 pass no initializer); transposed convs get the
    bilinear kernel (custom_layers.py:103).  `gain`/`bias_scale` let tests build
    "trained-like" nets whose logits are not degenerate."""
    params = {}
    for name, shape in fcn_param_shapes(prefix, cin, num_units, num_classes,
                                        batchnorm).items():
        leaf = name.split('/')[-1]
        layer = name.split('/')[-2]
        if layer in ('upscore_conv5', 'upscore') and leaf == 'kernel':
            params[name] = bilinear_filter(shape)
        elif leaf == 'kernel':
            fan_in = shape[0] * shape[1] * shape[2]
            fan_out = shape[0] * shape[1] * shape[3]
            limit = gain * np.sqrt(6.0 / (fan_in + fan_out))
            params[name] = rng.uniform(-limit, limit, size=shape).astype(np.float32)
        elif leaf == 'bias':
            params[name] = (bias_scale * rng.standard_normal(shape)).astype(np.float32)
        elif leaf in ('gamma', 'moving_variance'):
            params[name] = (1.0 + 0.1 * rng.random(shape)).astype(np.float32)
        else:
            params[name] = (0.1 * rng.standard_normal(shape)).astype(np.float32)
    return params


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def _batchnorm(x, params, scope):
    """tf.layers.batch_normalization at test time: gamma*(x-mean)/sqrt(var+1e-3)+beta
    (custom_layers.py:116,132-134; TF defaults epsilon=1e-3)."""
    g = params[scope + '/gamma']
    b = params[scope + '/beta']
    m = params[scope + '/moving_mean']
    v = params[scope + '/moving_variance']
    return (x - m) / np.sqrt(v + np.asarray(1e-3, x.dtype)) * g + b


def conv2d(x, params, scope, activation=True, batchnorm=False):
    """custom_layers.py:124-139: 'same' stride-1 conv + bias, [BN], [ReLU].
    x NHWC, kernel HWIO."""
    w = params[scope + '/kernel'].astype(x.dtype)
    b = params[scope + '/bias'].astype(x.dtype)
    pad = (w.shape[0] - 1) // 2
    y = F.conv2d(_t(x).permute(0, 3, 1, 2), _t(w).permute(3, 2, 0, 1), _t(b), padding=pad)
    y = y.permute(0, 2, 3, 1).contiguous().numpy()
    if batchnorm:
        y = _batchnorm(y, params, scope).astype(x.dtype)
    if activation:
        y = np.maximum(y, 0)
    return y


def deconv2d(x, params, scope, stride, activation=True, batchnorm=False):
    """custom_layers.py:71-121: conv2d_transpose 'same', no bias, kernel [kh,kw,Cout,Cin],
    then [BN], ReLU.  SAME padding for (k=4,s=2),(k=16,s=8) is (k-s)/2 each side and no
    spatial flip: w_torch[ci,co,y,x] = w_tf[y,x,co,ci]."""
    w = params[scope + '/kernel'].astype(x.dtype)
    k = w.shape[0]
    y = F.conv_transpose2d(_t(x).permute(0, 3, 1, 2), _t(w).permute(3, 2, 0, 1),
                           stride=stride, padding=(k - stride) // 2)
    y = y.permute(0, 2, 3, 1).contiguous().numpy()
    if batchnorm:
        y = _batchnorm(y, params, scope).astype(x.dtype)
    if activation:
        y = np.maximum(y, 0)
    return y


def max_pool2x2(x):
    """simple_fcn.py:41,44,48,58: max_pooling2d([2,2],[2,2]) 'valid'."""
    n, h, w, c = x.shape
    x = x[:, :h // 2 * 2, :w // 2 * 2]
    return x.reshape(n, h // 2, 2, w // 2, 2, c).max(axis=(2, 4))


def dropout(x, rate, keep_mask):
    """tf.layers.dropout(training=True) -> tf.nn.dropout: y = (x / keep_prob) * mask with
    mask = floor(keep_prob + U[0,1)).  `keep_mask` is that 0/1 mask
    (same shape as x), supplied by the caller so that device and oracle share it."""
    keep = np.asarray(1.0 - rate, x.dtype)
    return (x / keep) * keep_mask.astype(x.dtype)


def encoder(x, params, prefix, num_units, dropout_rate=0.0, dropout_layers=(),
            batchnorm=False, masks=None):
    """simple_fcn.py:10-87.  Replicates the reference quirk at :61 (pool4 dropout is gated
    by 'pool3' in dropout_layers)."""
    masks = masks or {}
    p = lambda s: prefix + '/' + s
    l = {}
    cur = x
    l['conv1_1'] = conv2d(cur, params, p('conv1_1'), batchnorm=batchnorm)
    l['conv1_2'] = conv2d(l['conv1_1'], params, p('conv1_2'), batchnorm=batchnorm)
    l['pool1'] = max_pool2x2(l['conv1_2'])
    l['conv2_1'] = conv2d(l['pool1'], params, p('conv2_1'), batchnorm=batchnorm)
    l['conv2_2'] = conv2d(l['conv2_1'], params, p('conv2_2'), batchnorm=batchnorm)
    l['pool2'] = max_pool2x2(l['conv2_2'])
    l['conv3_1'] = conv2d(l['pool2'], params, p('conv3_1'), batchnorm=batchnorm)
    l['conv3_2'] = conv2d(l['conv3_1'], params, p('conv3_2'), batchnorm=batchnorm)
    l['conv3_3'] = conv2d(l['conv3_2'], params, p('conv3_3'), batchnorm=batchnorm)
    l['pool3'] = max_pool2x2(l['conv3_3'])
    last = l['pool3']
    if 'pool3' in dropout_layers:
        l['pool3_drop'] = dropout(l['pool3'], dropout_rate, masks['pool3'])
        last = l['pool3_drop']
    l['conv4_1'] = conv2d(last, params, p('conv4_1'), batchnorm=batchnorm)
    l['conv4_2'] = conv2d(l['conv4_1'], params, p('conv4_2'), batchnorm=batchnorm)
    l['conv4_3'] = conv2d(l['conv4_2'], params, p('conv4_3'), batchnorm=batchnorm)
    l['pool4'] = max_pool2x2(l['conv4_3'])
    last = l['pool4']
    if 'pool3' in dropout_layers:  # sic, simple_fcn.py:61
        l['pool4_drop'] = dropout(l['pool4'], dropout_rate, masks['pool4'])
        last = l['pool4_drop']
    l['conv5_1'] = conv2d(last, params, p('conv5_1'), batchnorm=batchnorm)
    l['conv5_2'] = conv2d(l['conv5_1'], params, p('conv5_2'), batchnorm=batchnorm)
    l['conv5_3'] = conv2d(l['conv5_2'], params, p('conv5_3'), batchnorm=batchnorm)
    conv4_3 = l['conv4_3']
    if 'conv4_3' in dropout_layers:
        conv4_3 = dropout(conv4_3, dropout_rate, masks['conv4_3'])
    l['score_conv4'] = conv2d(conv4_3, params, p('score_conv4'), batchnorm=batchnorm)
    conv5_3 = l['conv5_3']
    if 'conv5_3' in dropout_layers:
        conv5_3 = dropout(conv5_3, dropout_rate, masks['conv5_3'])
    l['score_conv5'] = conv2d(conv5_3, params, p('score_conv5'), batchnorm=batchnorm)
    l['upscore_conv5'] = deconv2d(l['score_conv5'], params, p('upscore_conv5'), 2,
                                  batchnorm=batchnorm)
    l['fused'] = l['score_conv4'] + l['upscore_conv5']  # tf.add_n, simple_fcn.py:85
    return l


def decoder(features, params, prefix, num_units, num_classes, dropout_rate=None,
            batchnorm=False, masks=None):
    """simple_fcn.py:90-134."""
    masks = masks or {}
    if dropout_rate is not None:
        features = dropout(features, dropout_rate, masks['features'])
    up = deconv2d(features, params, prefix + '/upscore', 8, batchnorm=batchnorm)
    score = conv2d(up, params, prefix + '/score', activation=False, batchnorm=batchnorm)
    return {'upscore': up, 'score': score}


def fcn(x, params, prefix, num_units, num_classes, dropout_rate=0.0, dropout_layers=(),
        batchnorm=False, masks=None):
    """simple_fcn.py:137-170."""
    layers = encoder(x, params, prefix, num_units, dropout_rate, dropout_layers,
                     batchnorm, masks)
    layers.update(decoder(layers['fused'], params, prefix, num_units, num_classes,
                          dropout_rate=(dropout_rate if 'features' in dropout_layers
                                        else None),
                          batchnorm=batchnorm, masks=masks))
    return layers


def softmax(score):
    """tf.nn.softmax over the last axis (basic_fusion_model.py:21)."""
    m = score.max(axis=-1, keepdims=True)
    e = np.exp(score - m)
    return e / e.sum(axis=-1, keepdims=True)


def argmax_first(x, axis=-1):
    """tf.argmax: int64, first maximal index (numpy has the same tie rule)."""
    return np.argmax(x, axis=axis).astype(np.int64)


def test_pipeline(x, params, prefix, num_units, num_classes, **kw):
    """basic_fusion_model.py:9-23 for expert_model='fcn' (trainable=False, batchnorm=False)."""
    out = fcn(x, params, prefix, num_units, num_classes, batchnorm=False, **kw)
    out['prob'] = softmax(out['score'])
    out['classification'] = argmax_first(out['prob'])
    return out


def cross_entropy(score, labels, num_classes):
    """simple_fcn.py:205-215 + utils.py:43-53 + base_model.py:198-201:
    -sum(onehot * log_softmax(score)) / (1e-20 + sum(onehot)); labels outside [0,C) have an
    all-zero one-hot row."""
    m = score.max(axis=-1, keepdims=True)
    logp = score - m - np.log(np.exp(score - m).sum(axis=-1, keepdims=True))
    valid = (labels >= 0) & (labels < num_classes)
    idx = np.where(valid, labels, 0)
    picked = np.take_along_axis(logp, idx[..., None], axis=-1)[..., 0]
    return float(-(picked * valid).sum() / (1e-20 + valid.sum()))


def vgg16_tower(x, params, prefix):
    """xview/models/vgg16.py:7-51: un-scoped variable names `<prefix>_conv1_1/kernel`."""
    l = {}
    cur = x
    for name, _ in CONV_LAYERS:
        cur = conv2d(cur, params, '%s_%s' % (prefix, name))
        l[name] = cur
        if name in ('conv1_2', 'conv2_2', 'conv3_3', 'conv4_3'):
            cur = max_pool2x2(cur)
            l['pool' + name[4]] = cur
    return l


def fusion_fcn(inputs, params, prefixes, num_units, num_classes):
    """xview/models/fusion_fcn.py:11-40: two VGG16 towers, channel concat of conv4_3 / conv5_3,
    1x1 score convs, bilinear upscore + add, decoder with prefix 'fused'."""
    layers = {m: vgg16_tower(inputs[m], params, prefix) for m, prefix in prefixes.items()}
    layers['concat_conv4'] = np.concatenate([layers[m]['conv4_3'] for m in prefixes], axis=3)
    layers['concat_conv5'] = np.concatenate([layers[m]['conv5_3'] for m in prefixes], axis=3)
    layers['score_conv4'] = conv2d(layers['concat_conv4'], params, 'fused_score_conv4')
    layers['score_conv5'] = conv2d(layers['concat_conv5'], params, 'fused_score_conv5')
    layers['upscore_conv5'] = deconv2d(layers['score_conv5'], params, 'fused_upscore_conv5', 2)
    layers['features'] = layers['score_conv4'] + layers['upscore_conv5']
    # fusion_fcn.py:39 passes no batchnorm argument: the decoder's default batchnorm=True applies
    layers.update(decoder(layers['features'], params, 'fused', num_units, num_classes,
                          batchnorm=True))
    return layers


def fusion_fcn_params(prefixes, channels, num_units, num_classes, rng, gain=1.0, bias_scale=0.0):
    """Random init of fusion_fcn's variables under the names the reference gives them."""
    params = {}
    for m, prefix in prefixes.items():
        tower = glorot_fcn_params('x', channels[m], num_units, num_classes, rng, gain, bias_scale)
        for name, _ in CONV_LAYERS:
            for leaf in ('kernel', 'bias'):
                params['%s_%s/%s' % (prefix, name, leaf)] = tower['x/%s/%s' % (name, leaf)]
    cin = 512 * len(prefixes)
    for name in ('fused_score_conv4', 'fused_score_conv5'):
        limit = gain * np.sqrt(6.0 / (cin + num_units))
        params[name + '/kernel'] = rng.uniform(-limit, limit, size=(1, 1, cin, num_units)).astype(
            np.float32)
        params[name + '/bias'] = (bias_scale * rng.standard_normal(num_units)).astype(np.float32)
    params['fused_upscore_conv5/kernel'] = bilinear_filter((4, 4, num_units, num_units))
    params['fused/upscore/kernel'] = bilinear_filter((16, 16, num_units, num_units))
    limit = gain * np.sqrt(6.0 / (num_units + num_classes))
    params['fused/score/kernel'] = rng.uniform(-limit, limit, size=(1, 1, num_units, num_classes)
                                               ).astype(np.float32)
    params['fused/score/bias'] = (bias_scale * rng.standard_normal(num_classes)).astype(np.float32)
    # batch-norm variables of the decoder (non-identity statistics unless bias_scale == 0)
    for scope, channels in (('fused/upscore', num_units), ('fused/score', num_classes)):
        jitter = 1.0 if bias_scale else 0.0
        params[scope + '/gamma'] = (1 + 0.2 * jitter * rng.standard_normal(channels)).astype(
            np.float32)
        params[scope + '/beta'] = (0.1 * jitter * rng.standard_normal(channels)).astype(np.float32)
        params[scope + '/moving_mean'] = (0.1 * jitter * rng.standard_normal(channels)).astype(
            np.float32)
        params[scope + '/moving_variance'] = (1 + 0.3 * jitter * rng.random(channels)).astype(
            np.float32)
    return params
