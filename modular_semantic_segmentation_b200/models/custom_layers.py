"""Host-side counterparts of xview/models/custom_layers.py that the B200 path needs:
variable initialisers and thin functional wrappers over the device layers."""
import numpy as np

from .. import device as dev


def bilinear_filter_initializer(filter_shape):
    """custom_layers.py:8-25: [kh,kw,Cout,Cin] transposed-conv kernel holding a bilinear
    interpolation kernel on the channel diagonal (factor/centre derived from the width)."""
    width, height = filter_shape[0], filter_shape[1]
    factor = np.ceil(width / 2.0)
    center = (2 * factor - 1 - factor % 2) / (2.0 * factor)
    xs = 1 - np.abs(np.arange(width) / factor - center)
    ys = 1 - np.abs(np.arange(height) / factor - center)
    weights = np.zeros(filter_shape, dtype=np.float32)
    plane = np.outer(xs, ys).astype(np.float32)
    for i in range(filter_shape[2]):
        weights[:, :, i, i] = plane
    return weights


def glorot_uniform(shape, rng):
    """tf.layers default kernel initialiser (custom_layers.py:131,138 pass none)."""
    fan_in = shape[0] * shape[1] * shape[2]
    fan_out = shape[0] * shape[1] * shape[3]
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def softmax(inputs, temperature=1):
    """custom_layers.py:236-245: temperature-scaled softmax over the last axis of a CUDA float32
    tensor, on the device (xv_softmax_argmax)."""
    scaled = inputs / temperature if temperature != 1 else inputs
    return dev.softmax_argmax(scaled.contiguous(), want_prob=True)[0]


def log_softmax(inputs, num_classes=None):
    """custom_layers.py:222-233 (the DA-RNN form d - log(sum(exp(d))), d = x - max(x)): computed
    from the device softmax as log(p) where p > 0, exact in the limit the reference takes."""
    import torch
    prob = softmax(inputs)
    shifted = inputs - inputs.max(dim=-1, keepdim=True).values
    return torch.where(prob > 0, torch.log(prob), shifted)


def entropy(x, axis=-1):
    """custom_layers.py:251-256: normed entropy -sum(x log clip(x, 1e-10, 1)) / log(C) of CUDA
    float32 probabilities, through the MC-moments kernel (one "sample": the entropy of its
    mean is the entropy of x)."""
    if axis not in (-1, x.dim() - 1):
        x = x.movedim(axis, -1)
    return dev.mc_moments(x.contiguous().unsqueeze(0), want=('entropy',))['entropy']


def conv2d(inputs, kernel, bias=None, activation=True, precision='bf16'):
    """custom_layers.py:124-139 (no batch norm): CUDA float32 NHWC in / out."""
    return dev.conv2d(inputs, kernel, bias, relu=activation, precision=precision)


def deconv2d(inputs, kernel, strides, activation=True):
    """custom_layers.py:71-121 (no batch norm)."""
    return dev.deconv2d(inputs, kernel, strides, relu=activation)
