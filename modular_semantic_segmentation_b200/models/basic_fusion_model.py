"""Mirror of xview/models/basic_fusion_model.py: one FCN expert per modality + a fusion rule."""
import numpy as np
import torch

from .base_model import BaseModel
from .adapnet import build_adapnet
from .simple_fcn import build_expert


def build_test_pipeline(prefix, channels, **config):
    """basic_fusion_model.py:9-23: the expert network behind `test_pipeline`, 'adapnet' or
    'fcn'."""
    if config['expert_model'] == 'adapnet':
        return build_adapnet(prefix, channels, config['num_units'], config['num_classes'],
                             precision=config.get('precision', 'bf16'),
                             rng=np.random.default_rng(config.get('seed')))
    if config['expert_model'] == 'fcn':
        return build_expert(prefix, channels, config['num_units'], config['num_classes'],
                            batchnorm=False, precision=config.get('precision', 'bf16'),
                            rng=np.random.default_rng(config.get('seed')))
    raise UserWarning('ERROR: Expert Model %s not found' % config['expert_model'])


class FusionModel(BaseModel):
    """basic_fusion_model.py:26-66.  Subclasses implement `_fusion(expert_outputs, fetch)`."""

    # what `_fusion` needs from each expert: any of 'label', 'prob'
    expert_wants = ('label',)

    def __init__(self, name=None, output_dir=None, **config):
        self.modalities = list(config['prefixes'].keys())
        BaseModel.__init__(self, name=name, output_dir=output_dir, custom_training=True,
                           **config)

    def _expert_prefix(self, modality):
        return self.config['prefixes'][modality]

    def _build_graph(self):
        for m in self.modalities:
            channels = self.config['num_channels'][m]
            expert, variables = build_test_pipeline(self._expert_prefix(m), channels,
                                                    **self.config)
            self._register_expert(self._expert_prefix(m), expert, variables)
        self.prediction = 'prediction'

    def _arrival_order(self, batch):
        """Modalities in the order their arrays reach the device (smallest first, see
        base_model.upload_order): the expert of the first one starts while the others are still
        being copied.  The fusion rules index experts by `self.modalities`, so the order in
        which the experts RUN does not matter for the result."""
        from .base_model import _nbytes
        return sorted(self.modalities, key=lambda m: _nbytes(dict.__getitem__(batch, m))
                      if isinstance(batch, dict) else 0)

    def _expert_outputs(self, batch, wants, label_dtype):
        def run(m, x):
            out = self._experts[self._expert_prefix(m)].forward(x, want=wants,
                                                                label_dtype=label_dtype)
            if 'label' in out:
                out['classification'] = out['label']
            return out
        return self._run_experts(batch, run)

    def _fusion(self, expert_outputs, fetch, label_dtype):
        raise NotImplementedError

    def _run_batch(self, batch, fetch='prediction'):
        label_dtype = torch.uint8 if fetch == 'prediction_compact' else torch.int64
        self.expert_outputs = self._expert_outputs(batch, self.expert_wants, label_dtype)
        return self._fusion(self.expert_outputs, fetch, label_dtype)
