"""Mirror of xview/models/average_mix.py."""
from .. import device as dev
from .basic_fusion_model import FusionModel


class AverageFusion(FusionModel):
    """average_mix.py:6-21: argmax of the mean of the experts' probabilities."""

    output_attrs = ('prediction', 'fused_score')
    expert_wants = ('prob',)

    def __init__(self, output_dir=None, **config):
        FusionModel.__init__(self, name='AverageFusion', output_dir=output_dir, **config)

    def _fusion(self, expert_outputs, fetch, label_dtype):
        probs = [expert_outputs[m]['prob'] for m in self.modalities]
        score, label = dev.average_fuse(probs, want_score=(fetch == 'fused_score'),
                                        label_dtype=label_dtype)
        return score if fetch == 'fused_score' else label
