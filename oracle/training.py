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


def _conv(x, p, scope, relu=True):
    w = p[scope + '/kernel']
    y = F.conv2d(x, w.permute(3, 2, 0, 1), p[scope + '/bias'], padding=(w.shape[0] - 1) // 2)
    return F.relu(y) if relu else y


def _deconv(x, p, scope, stride):
    w = p[scope + '/kernel']
    k = w.shape[0]
    return F.relu(F.conv_transpose2d(x, w.permute(3, 2, 0, 1), stride=stride,
                                     padding=(k - stride) // 2))


def loss_and_grads(params, prefix, x, labels, num_classes, dtype=torch.float32,
                   train_encoder=True):
    """Returns (loss, {variable name: gradient}) for one batch.  x NHWC, labels [N,H,W] int;
    labels outside [0, C) contribute nothing (all-zero one-hot row)."""
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
        h = _conv(h, p, s(name))
        acts[name] = h
        if name in ('conv1_2', 'conv2_2', 'conv3_3', 'conv4_3'):
            h = F.max_pool2d(h, 2)
    score4 = _conv(acts['conv4_3'], p, s('score_conv4'))
    score5 = _conv(acts['conv5_3'], p, s('score_conv5'))
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
