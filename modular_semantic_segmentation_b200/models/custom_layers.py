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


def conv2d(inputs, kernel, bias=None, activation=True, precision='bf16'):
    """custom_layers.py:124-139 (no batch norm): CUDA float32 NHWC in / out."""
    return dev.conv2d(inputs, kernel, bias, relu=activation, precision=precision)


def deconv2d(inputs, kernel, strides, activation=True):
    """custom_layers.py:71-121 (no batch norm)."""
    return dev.deconv2d(inputs, kernel, strides, relu=activation)
