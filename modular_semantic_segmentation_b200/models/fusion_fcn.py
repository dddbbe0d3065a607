"""Mirror of xview/models/fusion_fcn.py + vgg16.py: the mid-level fusion network (two VGG16
towers, channel concat of conv4_3 / conv5_3, one shared head and decoder)."""
from collections import OrderedDict

import numpy as np
import torch

from .. import device as dev
from .base_model import BaseModel
from .custom_layers import bilinear_filter_initializer, glorot_uniform
from .simple_fcn import CONV_LAYERS


def init_fusion_fcn_variables(prefixes, num_channels, num_units, num_classes, rng=None):
    """Variables under the reference's names: `<prefix>_conv1_1/kernel` for the towers
    (vgg16.py:17 - no variable scope), `fused_score_conv4/5`, `fused_upscore_conv5`
    (fusion_fcn.py:28-35) and `fused/upscore`, `fused/score` from the decoder (:39)."""
    rng = rng if rng is not None else np.random.default_rng()
    v = OrderedDict()
    for modality, prefix in prefixes.items():
        cin = num_channels[modality]
        for name, cout in CONV_LAYERS:
            v['%s_%s/kernel' % (prefix, name)] = glorot_uniform((3, 3, cin, cout), rng)
            v['%s_%s/bias' % (prefix, name)] = np.zeros(cout, np.float32)
            cin = cout
    cat = 512 * len(prefixes)
    for name in ('fused_score_conv4', 'fused_score_conv5'):
        v[name + '/kernel'] = glorot_uniform((1, 1, cat, num_units), rng)
        v[name + '/bias'] = np.zeros(num_units, np.float32)
    v['fused_upscore_conv5/kernel'] = bilinear_filter_initializer((4, 4, num_units, num_units))
    v['fused/upscore/kernel'] = bilinear_filter_initializer((16, 16, num_units, num_units))
    v['fused/score/kernel'] = glorot_uniform((1, 1, num_units, num_classes), rng)
    v['fused/score/bias'] = np.zeros(num_classes, np.float32)
    # fusion_fcn.py:39 calls decoder() without a batchnorm argument and decoder defaults to
    # batchnorm=True (simple_fcn.py:91): `fused/upscore` is deconv -> BN -> ReLU and
    # `fused/score` is conv -> BN, each with the four tf.layers.batch_normalization variables
    for scope, channels in (('fused/upscore', num_units), ('fused/score', num_classes)):
        v[scope + '/gamma'] = np.ones(channels, np.float32)
        v[scope + '/beta'] = np.zeros(channels, np.float32)
        v[scope + '/moving_mean'] = np.zeros(channels, np.float32)
        v[scope + '/moving_variance'] = np.ones(channels, np.float32)
    return v


class _FusionFcnDevice(object):
    """The device side: one encoder handle per modality + one head handle."""

    def __init__(self, prefixes, num_channels, num_units, num_classes, precision='bf16'):
        self.prefixes = OrderedDict(prefixes)
        self.towers = OrderedDict(
            (m, dev.FcnExpert(num_channels[m], num_units, num_classes, precision=precision,
                              role='encoder')) for m in self.prefixes)
        self.head = dev.FcnExpert(1, num_units, num_classes, precision=precision, role='head',
                                  head_cin=512 * len(self.prefixes), batchnorm='decoder')

    def set_variables(self, variables):
        for m, prefix in self.prefixes.items():
            head = prefix + '_'
            self.towers[m].set_params({n[len(head):]: a for n, a in variables.items()
                                       if n.startswith(head) and '/' in n and
                                       n[len(head):].startswith('conv')})
        self.head.set_params({
            'score_conv4/kernel': variables['fused_score_conv4/kernel'],
            'score_conv4/bias': variables['fused_score_conv4/bias'],
            'score_conv5/kernel': variables['fused_score_conv5/kernel'],
            'score_conv5/bias': variables['fused_score_conv5/bias'],
            'upscore_conv5/kernel': variables['fused_upscore_conv5/kernel'],
            'upscore/kernel': variables['fused/upscore/kernel'],
            'score/kernel': variables['fused/score/kernel'],
            'score/bias': variables['fused/score/bias']})
        self.head.set_params({'%s/%s' % (scope, leaf): variables['fused/%s/%s' % (scope, leaf)]
                              for scope in ('upscore', 'score')
                              for leaf in ('gamma', 'beta', 'moving_mean', 'moving_variance')})

    def forward(self, inputs, want=('label',), label_dtype=torch.int64):
        for m in self.prefixes:
            self.towers[m].forward_encoder(inputs[m])
        out = self.head.forward_head(list(self.towers.values()), want=want,
                                     label_dtype=label_dtype)
        if 'label' in out:
            out['classification'] = out['label']
        return out

    def close(self):
        for t in self.towers.values():
            t.close()
        self.head.close()


_STORE = {}


def fusion_fcn(inputs, prefixes, num_units, num_classes, trainable=True, is_training=False,
               reuse=False, precision='bf16', want=('score',), params=None):
    """fusion_fcn.py:11-40 as called by experiments/timing.py:29-31.  inputs: {modality: CUDA
    float32 [N,H,W,C]}.  Variables are created on first use (random init as in the reference)."""
    channels = {m: int(inputs[m].shape[-1]) for m in prefixes}
    key = (tuple(prefixes.items()), tuple(channels.items()), num_units, num_classes, precision)
    if key not in _STORE:
        net = _FusionFcnDevice(prefixes, channels, num_units, num_classes, precision)
        variables = init_fusion_fcn_variables(prefixes, channels, num_units, num_classes)
        net.set_variables(variables)
        _STORE[key] = (net, variables)
    net, variables = _STORE[key]
    if params is not None:
        variables.update(params)
        net.set_variables(variables)
    return net.forward(inputs, want=want)


class FusionFCN(BaseModel):
    """fusion_fcn.py:43-121.  The reference class uses a pre-refactor BaseModel signature and
    cannot be constructed as shipped; this class offers the same
    network behind the current BaseModel surface (predict / score / import_weights)."""

    output_attrs = ('prediction', 'prob', 'score')

    def __init__(self, data_description, prefixes, num_channels, num_units, output_dir=None,
                 **config):
        self.modalities = list(prefixes.keys())
        BaseModel.__init__(self, data_description, name='FusionFCN', output_dir=output_dir,
                           custom_training=True, prefixes=prefixes, num_channels=num_channels,
                           num_units=num_units, **config)

    def _build_graph(self):
        cfg = self.config
        self._net = _FusionFcnDevice(cfg['prefixes'], cfg['num_channels'], cfg['num_units'],
                                     cfg['num_classes'], cfg.get('precision', 'bf16'))
        self.variables.update(init_fusion_fcn_variables(
            cfg['prefixes'], cfg['num_channels'], cfg['num_units'], cfg['num_classes'],
            rng=np.random.default_rng(cfg.get('seed'))))
        self._push_variables()
        self.prediction = 'prediction'

    def _push_variables(self, prefix=None):
        self._net.set_variables(self.variables)

    def _run_batch(self, batch, fetch='prediction'):
        if fetch == 'prediction':
            return self._net.forward(batch, want=('label',))['label']
        if fetch == 'prediction_compact':
            return self._net.forward(batch, want=('label',), label_dtype=torch.uint8)['label']
        return self._net.forward(batch, want=(fetch,))[fetch]

    def close(self):
        self._net.close()
        BaseModel.close(self)
