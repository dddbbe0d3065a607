"""Mirror of xview/models/adapnet.py: the Adapnet expert (ResNet-50 style blocks with atrous
stage-2 convolutions), functional `adapnet()` and the `Adapnet` model class, test-time graph on
the B200 kernels (batch norm on its moving statistics, folded into the convolutions)."""
from collections import OrderedDict

import numpy as np
import torch

from .. import device as dev
from .base_model import BaseModel
from .custom_layers import bilinear_filter_initializer, glorot_uniform

# (scope, kind, arguments) of adapnet.py:124-153
#   'a': (intermediate filters, filters, strides, shortcut_conv)
#   'b': (filters_1, filters_2, filters_3, dilation1, dilation2, shortcut_conv)
BLOCKS = [
    ('block_layer_1', 'a', (64, 256, 1, True)), ('block_layer_2', 'a', (64, 256, 1, False)),
    ('block_layer_3', 'a', (64, 256, 1, False)), ('block_layer_4', 'a', (128, 512, 2, True)),
    ('block_layer_5', 'a', (128, 512, 1, False)), ('block_layer_6', 'a', (128, 512, 1, False)),
    ('block_layer_7', 'b', (128, 64, 512, 1, 2, False)),
    ('block_layer_8', 'a', (256, 1024, 2, True)), ('block_layer_9', 'a', (256, 1024, 1, False)),
    ('block_layer_10', 'b', (256, 256, 1024, 1, 2, False)),
    ('block_layer_11', 'b', (256, 256, 1024, 1, 4, False)),
    ('block_layer_12', 'b', (256, 256, 1024, 1, 8, False)),
    ('block_layer_13', 'b', (256, 256, 1024, 1, 16, False)),
    ('block_layer_14', 'b', (512, 512, 2048, 2, 4, True)),
    ('block_layer_15', 'b', (512, 512, 2048, 2, 8, False)),
    ('block_layer_16', 'b', (512, 512, 2048, 2, 16, False)),
]

LAYER_NAMES = (['block_0_1', 'block_0_2', 'block_0_pool'] +
               ['block_%d' % i for i in range(1, 17)] + ['shortcut', 'merge', 'score'])


def init_adapnet_variables(prefix, num_channels, num_units, num_classes, rng=None):
    """Variables of one Adapnet expert under their TensorFlow names with the values
    `tf.global_variables_initializer()` gives them: Glorot-uniform kernels, zero biases (only
    where use_bias is left at its default: block_0_*, shortcut, first_deconvolution_conv),
    bilinear transposed-conv kernels [k,k,Cout,Cin] and identity batch-norm statistics
    (adapnet.py:35-36,78-79,116-166; custom_layers.py:92,103,114-116,132-134)."""
    rng = rng if rng is not None else np.random.default_rng()
    v = OrderedDict()

    def bn(scope, c):
        v[scope + '/gamma'] = np.ones(c, np.float32)
        v[scope + '/beta'] = np.zeros(c, np.float32)
        v[scope + '/moving_mean'] = np.zeros(c, np.float32)
        v[scope + '/moving_variance'] = np.ones(c, np.float32)

    def conv(scope, k, cin, cout, bias=False):
        scope = '%s/%s' % (prefix, scope)
        v[scope + '/kernel'] = glorot_uniform((k, k, cin, cout), rng)
        if bias:
            v[scope + '/bias'] = np.zeros(cout, np.float32)
        bn(scope, cout)

    conv('block_0_1', 3, num_channels, 64, bias=True)
    conv('block_0_2', 7, 64, 64, bias=True)
    c = 64
    for scope, kind, args in BLOCKS:
        if kind == 'a':
            mid, out, _, shortcut = args
            conv(scope + '/stage_1', 1, c, mid)
            conv(scope + '/stage_2', 3, mid, mid)
            conv(scope + '/stage_3', 1, mid, out)
        else:
            f1, f2, out, _, _, shortcut = args
            conv(scope + '/stage_1', 1, c, f1)
            conv(scope + '/stage_2_1', 3, f1, f2 // 2)
            conv(scope + '/stage_2_2', 3, f1, f2 // 2)
            conv(scope + '/stage_3', 1, f2, out)
        if shortcut:
            conv(scope + '/shortcut', 1, c, out)
        c = out
        if scope == 'block_layer_7':
            conv('shortcut', 1, c, num_units, bias=True)
    conv('first_deconvolution_conv', 1, 2048, 2048, bias=True)
    for scope, k, cout, cin in (('first_deconvolution_upconv', 4, num_units, 2048),
                                ('second_deconvolution_upconv', 16, num_classes, num_units)):
        scope = '%s/%s' % (prefix, scope)
        v[scope + '/kernel'] = bilinear_filter_initializer((k, k, cout, cin))
        bn(scope, cout)
    return v


def build_adapnet(prefix, num_channels, num_units, num_classes, precision='bf16', rng=None):
    """Creates the device expert plus its freshly initialised variables."""
    variables = init_adapnet_variables(prefix, num_channels, num_units, num_classes, rng)
    expert = dev.FcnExpert(num_channels, num_units, num_classes, precision=precision,
                           arch='adapnet')
    return expert, variables


def adapnet(inputs, prefix, num_units, num_classes, is_training=False, reuse=None, params=None,
            precision='bf16', layers=('score',), seed=0):
    """adapnet.py:99-173 as called by the fusion models: runs the test-time graph on `inputs`
    (float32 [N,H,W,cin], numpy or CUDA tensor) and returns {'score': CUDA tensor} plus any
    further block outputs named in `layers` (numpy, via the layer inspection call).  `params`:
    dict of variables named '<prefix>/<scope>/<var>'; missing -> fresh initial values."""
    if is_training:
        raise UserWarning('ERROR: only the test-time Adapnet graph runs on the B200 path')
    x = dev.to_device(inputs, torch.float32)
    expert, variables = build_adapnet(prefix, x.shape[-1], num_units, num_classes, precision,
                                      np.random.default_rng(seed))
    try:
        if params is not None:
            variables.update({k: v for k, v in params.items() if k in variables})
        expert.set_params({name[len(prefix) + 1:]: value for name, value in variables.items()})
        out = expert.forward(x, want=('score',))
        for name in layers:
            if name != 'score':
                out[name] = expert.layer(name)
        return out
    finally:
        expert.close()


class Adapnet(BaseModel):
    """adapnet.py:176-235.  Test-time network: softmax + argmax of the Adapnet scores."""

    output_attrs = ('prediction', 'prob', 'score')

    def __init__(self, data_description, prefix=None, output_dir=None, **config):
        standard_config = {'train_encoder': True}
        standard_config.update(config)
        self.prefix = config['modality'] if prefix is None else prefix
        self.modality = config['modality']
        BaseModel.__init__(self, data_description, output_dir=output_dir, **standard_config)

    def _build_graph(self):
        channels = self.testdata_description[1][self.modality][-1]
        expert, variables = build_adapnet(
            self.prefix, channels, self.config['num_units'], self.config['num_classes'],
            precision=self.config.get('precision', 'bf16'),
            rng=np.random.default_rng(self.config.get('seed')))
        self._register_expert(self.prefix, expert, variables)
        self.prediction = 'prediction'

    def fit(self, *args, **kwargs):
        raise UserWarning('ERROR: Adapnet training (batch-norm training mode) is not on the B200 '
                          'path; import trained weights with import_weights()')

    def _run_batch(self, batch, fetch='prediction'):
        expert = self._experts[self.prefix]
        x = batch[self.modality]
        if fetch == 'prediction':
            return expert.forward(x, want=('label',))['label']
        if fetch == 'prediction_compact':
            return expert.forward(x, want=('label',), label_dtype=torch.uint8)['label']
        return expert.forward(x, want=(fetch,))[fetch]
