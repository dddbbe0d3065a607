"""Oracle: one training step of the SimpleFCN expert (TEST INFRASTRUCTURE ONLY).

Restates the training branch of xview/models/simple_fcn.py:201-215 (fcn -> log_softmax ->
cross_entropy of xview/models/utils.py:43-53 on one-hot labels, base_model.py:198-201) and the
optimizer of base_model.py:153-162 (tf.train.AdamOptimizer defaults: beta1 0.9, beta2 0.999,
epsilon 1e-8, lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)).  Gradients come from torch
autograd on the CPU in float32 (or float64 for tolerance studies).  Batch normalisation and the
non-trainable transposed convolutions follow the reference: deconv kernels get no gradient.
PARITY UNPINNED at the TensorFlow boundary (see oracle/__init__).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .fcn import CONV_LAYERS


class _Bf16Both(torch.autograd.Function):
    """Value rounded to bfloat16 on the way forward, gradient rounded to bfloat16 on the way
    back: what happens to an activation the device stores in bf16 (and to the data gradient it
    stores in bf16 for the layer below)."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().to(g.dtype)


def _bf16_forward_only(w):
    """bf16 operand copy of an fp32 master weight: rounded in the forward and data-gradient
    passes, while its own gradient stays fp32 (straight-through)."""
    return w + (w.detach().bfloat16().to(w.dtype) - w.detach())


def _conv(x, p, scope, relu=True, emulate_bf16=False, store_bf16=True):
    w = p[scope + '/kernel']
    if emulate_bf16:
        w = _bf16_forward_only(w)
    y = F.conv2d(x, w.permute(3, 2, 0, 1), p[scope + '/bias'], padding=(w.shape[0] - 1) // 2)
    y = F.relu(y) if relu else y
    if emulate_bf16 and store_bf16:
        y = _Bf16Both.apply(y)
    return y


def _deconv(x, p, scope, stride):
    w = p[scope + '/kernel']
    k = w.shape[0]
    return F.relu(F.conv_transpose2d(x, w.permute(3, 2, 0, 1), stride=stride,
                                     padding=(k - stride) // 2))


def loss_and_grads(params, prefix, x, labels, num_classes, dtype=torch.float32,
                   train_encoder=True, emulate_bf16=False):
    """Returns (loss, {variable name: gradient}) for one batch.  x NHWC, labels [N,H,W] int;
    labels outside [0, C) contribute nothing (all-zero one-hot row).
    emulate_bf16: same graph with the device's storage precision restated - bf16 operand copies
    of the weights, encoder activations and their data gradients rounded to bf16, fp32 heads and
    decoder - so that the device's backward pass can be checked against
    autograd much more tightly than against the pure fp32 graph."""
    p = {}
    for name, value in params.items():
        t = torch.tensor(np.asarray(value), dtype=dtype)
        leaf = name.split('/')[-2]
        trainable = leaf not in ('upscore_conv5', 'upscore')
        if not train_encoder and leaf.startswith('conv'):
            trainable = False
        t.requires_grad_(trainable)
        p[name] = t
    s = lambda n: prefix + '/' + n
    h = torch.tensor(np.asarray(x), dtype=dtype).permute(0, 3, 1, 2)
    acts = {}
    for name, _ in CONV_LAYERS:
        h = _conv(h, p, s(name), emulate_bf16=emulate_bf16)
        acts[name] = h
        if name in ('conv1_2', 'conv2_2', 'conv3_3', 'conv4_3'):
            h = F.max_pool2d(h, 2)
    score4 = _conv(acts['conv4_3'], p, s('score_conv4'), emulate_bf16=emulate_bf16,
                   store_bf16=False)
    score5 = _conv(acts['conv5_3'], p, s('score_conv5'), emulate_bf16=emulate_bf16,
                   store_bf16=False)
    fused = score4 + _deconv(score5, p, s('upscore_conv5'), 2)
    up = _deconv(fused, p, s('upscore'), 8)
    score = _conv(up, p, s('score'), relu=False).permute(0, 2, 3, 1)      # NHWC
    logp = F.log_softmax(score, dim=-1)
    lab = torch.tensor(np.asarray(labels), dtype=torch.int64)
    valid = (lab >= 0) & (lab < num_classes)
    picked = torch.gather(logp, -1, lab.clamp(0, num_classes - 1)[..., None])[..., 0]
    loss = -(picked * valid).sum() / (1e-20 + valid.sum())
    loss.backward()
    grads = {n: t.grad.numpy() for n, t in p.items() if t.requires_grad and t.grad is not None}
    return float(loss.detach()), grads


def adam_update(params, grads, m, v, step, learning_rate=1e-4, beta1=0.9, beta2=0.999,
                epsilon=1e-8):
    """tf.train.AdamOptimizer.apply_gradients for step `step` (1-based); updates in place."""
    lr_t = learning_rate * np.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    for name, g in grads.items():
        m[name] = beta1 * m.get(name, 0.0) + (1 - beta1) * g
        v[name] = beta2 * v.get(name, 0.0) + (1 - beta2) * g * g
        params[name] = (params[name] - lr_t * m[name] / (np.sqrt(v[name]) + epsilon)).astype(
            params[name].dtype)
    return params


def adagrad_update(params, grads, acc, learning_rate=1e-4, initial_accumulator_value=0.1):
    """tf.train.AdagradOptimizer (base_model.py:157): acc starts at 0.1; acc += g^2;
    w -= lr * g / sqrt(acc).  Updates in place."""
    for name, g in grads.items():
        acc[name] = acc.get(name, initial_accumulator_value) + g * g
        params[name] = (params[name] - learning_rate * g / np.sqrt(acc[name])).astype(
            params[name].dtype)
    return params


def rmsprop_update(params, grads, ms, mom, learning_rate=1e-4, decay=0.9, momentum=0.0,
                   epsilon=1e-10):
    """tf.train.RMSPropOptimizer (base_model.py:159), TF 1.x defaults; the mean-square slot
    starts at ONE: ms = decay*ms + (1-decay)*g^2; mom = momentum*mom + lr*g/sqrt(ms+eps);
    w -= mom.  Updates in place."""
    for name, g in grads.items():
        ms[name] = decay * ms.get(name, 1.0) + (1 - decay) * g * g
        mom[name] = momentum * mom.get(name, 0.0) + learning_rate * g / np.sqrt(ms[name] + epsilon)
        params[name] = (params[name] - mom[name]).astype(params[name].dtype)
    return params



# ------------------------------------------------------------------ batch norm in training mode
BN_EPS = 1e-3        # tf.layers.batch_normalization defaults
BN_MOMENTUM = 0.99
BN_SCOPES = [n for n, _ in CONV_LAYERS] + ['score_conv4', 'score_conv5', 'upscore_conv5',
                                          'upscore', 'score']


def _bn_train(z, p, scope, stats):
    """tf.layers.batch_normalization(training=True) on an NCHW tensor: normalise with the
    statistics of the batch (biased variance), y = gamma * (z - mean) / sqrt(var + 1e-3) + beta;
    the batch mean / biased variance are recorded for the moving-average update."""
    mean = z.mean(dim=(0, 2, 3))
    var = z.var(dim=(0, 2, 3), unbiased=False)
    stats[scope] = (mean.detach().numpy().copy(), var.detach().numpy().copy(),
                    z.shape[0] * z.shape[2] * z.shape[3])
    zhat = (z - mean[None, :, None, None]) * torch.rsqrt(var + BN_EPS)[None, :, None, None]
    return zhat * p[scope + '/gamma'][None, :, None, None] + p[scope + '/beta'][None, :, None, None]


def loss_and_grads_bn(params, prefix, x, labels, num_classes, dtype=torch.float32,
                      emulate_bf16=False):
    """Training graph of simple_fcn.py:201-215 with batchnorm=True, is_training=True: every
    conv / transposed conv is followed by batch normalisation on batch statistics, then ReLU
    (custom_layers.py:112-119,127-136); the final score conv has batch norm but no activation.
    Returns (loss, gradients incl. gamma / beta, {scope: (batch mean, biased batch variance,
    count)}).  emulate_bf16: the device's storage precision restated (bf16 operand copies of the
    encoder / head weights, encoder pre-norm outputs z and activations y and their gradients
    rounded to bf16; fp32 heads and decoder)."""
    p = {}
    for name, value in params.items():
        t = torch.tensor(np.asarray(value), dtype=dtype)
        parts = name.split('/')
        trainable = not (parts[-2] in ('upscore_conv5', 'upscore') and parts[-1] == 'kernel')
        trainable = trainable and parts[-1] not in ('moving_mean', 'moving_variance')
        t.requires_grad_(trainable)
        p[name] = t
    s = lambda n: prefix + '/' + n
    stats = {}

    def conv_bn(h, scope, relu=True, q_weights=False, q_acts=False):
        w = p[s(scope) + '/kernel']
        if emulate_bf16 and q_weights:
            w = _bf16_forward_only(w)
        z = F.conv2d(h, w.permute(3, 2, 0, 1), p[s(scope) + '/bias'],
                     padding=(w.shape[0] - 1) // 2)
        if emulate_bf16 and q_acts:
            z = _Bf16Both.apply(z)
        y = _bn_train(z, p, s(scope), stats)
        y = F.relu(y) if relu else y
        if emulate_bf16 and q_acts:
            y = _Bf16Both.apply(y)
        return y

    def deconv_bn(h, scope, stride):
        w = p[s(scope) + '/kernel']
        k = w.shape[0]
        z = F.conv_transpose2d(h, w.permute(3, 2, 0, 1), stride=stride,
                               padding=(k - stride) // 2)
        return F.relu(_bn_train(z, p, s(scope), stats))

    h = torch.tensor(np.asarray(x), dtype=dtype).permute(0, 3, 1, 2)
    acts = {}
    for name, _ in CONV_LAYERS:
        h = conv_bn(h, name, q_weights=True, q_acts=True)
        acts[name] = h
        if name in ('conv1_2', 'conv2_2', 'conv3_3', 'conv4_3'):
            h = F.max_pool2d(h, 2)
    score4 = conv_bn(acts['conv4_3'], 'score_conv4', q_weights=True)
    score5 = conv_bn(acts['conv5_3'], 'score_conv5', q_weights=True)
    fused = score4 + deconv_bn(score5, 'upscore_conv5', 2)
    up = deconv_bn(fused, 'upscore', 8)
    score = conv_bn(up, 'score', relu=False).permute(0, 2, 3, 1)
    logp = F.log_softmax(score, dim=-1)
    lab = torch.tensor(np.asarray(labels), dtype=torch.int64)
    valid = (lab >= 0) & (lab < num_classes)
    picked = torch.gather(logp, -1, lab.clamp(0, num_classes - 1)[..., None])[..., 0]
    loss = -(picked * valid).sum() / (1e-20 + valid.sum())
    loss.backward()
    grads = {n: t.grad.numpy() for n, t in p.items() if t.requires_grad and t.grad is not None}
    stats = {k[len(prefix) + 1:]: v for k, v in stats.items()}
    return float(loss.detach()), grads, stats


def moving_average_update(moving_mean, moving_variance, batch_mean, batch_variance, count,
                          momentum=BN_MOMENTUM):
    """Update ops of tf.layers.batch_normalization (base_model.py:155-156 runs them with every
    training step): moving = momentum * moving + (1 - momentum) * batch statistic.  The fused TF
    kernel feeds the UNBIASED batch variance (Bessel-corrected) into the moving variance;
    documented assumption (no TensorFlow in the image)."""
    unbiased = batch_variance * (count / max(count - 1, 1))
    return (momentum * moving_mean + (1 - momentum) * batch_mean,
            momentum * moving_variance + (1 - momentum) * unbiased)
