"""Mirror of xview/models/variance_mix.py.  The reference class cannot be constructed as
shipped (pre-refactor BaseModel signature); the working statement is
experiments/timing.py:180-233, which this class follows."""
from .. import device as dev
from .. import sharding
from .basic_fusion_model import FusionModel


def variance_fusion(probs, variances, want_score=True):
    """variance_mix.py:7-15 on the device: inverse-variance weighted mean of the experts'
    probabilities.  variances: per-pixel float32 CUDA maps [N,H,W]."""
    return dev.variance_fuse(probs, variances, want_score=want_score)[0 if want_score else 1]


def mc_dropout_seed(model, stream_index):
    """Philox key of one MC-dropout call.  The reference draws fresh masks in every sess.run;
    here every `_run_batch` call advances a per-model counter that is folded into the seed
    (config `deterministic_dropout=True` freezes the masks, for tests)."""
    calls = getattr(model, '_mc_calls', 0)
    if stream_index == 0 and not model.config.get('deterministic_dropout', False):
        calls += 1
        model._mc_calls = calls
    base = int(model.config.get('seed') or 0)
    return (base * 1000003 + calls * 8191 + stream_index) & 0xFFFFFFFFFFFFFFFF


def split_samples_over_ranks(model):
    """(samples of this rank, rank) if the model is configured to split its MC-dropout SAMPLES
    over the ranks instead of the images (`split_samples=True`, `shard_images=False`: every rank
    sees every image - the batch-1 latency mode of SURVEY.md section 8e), else None.  Every rank
    needs at least two samples (moments of a single sample are degenerate)."""
    if not model.config.get('split_samples', False):
        return None
    rank, world = sharding.rank_world()
    if world == 1:
        return None
    if model.config.get('shard_images', True):
        raise UserWarning('ERROR: split_samples=True needs shard_images=False')
    mine = sharding.samples_for_rank(int(model.config['num_samples']), rank, world)
    if mine < 2:
        raise UserWarning('ERROR: split_samples needs at least 2 MC samples per rank')
    return mine, rank


class VarianceFusion(FusionModel):
    """variance_mix.py:18-83: MC-dropout (dropout after pool3, `num_samples` samples sharing
    one weight load) gives the per-pixel variance, a dropout-free pass gives the
    probabilities."""

    output_attrs = ('prediction', 'fused_score')

    def __init__(self, output_dir=None, **config):
        standard_config = {'learning_rate': 0.0}
        standard_config.update(config)
        if 'prefixes' not in standard_config:
            standard_config['prefixes'] = {m: m for m in standard_config['modalities']}
        FusionModel.__init__(self, 'VarianceMixture', output_dir=output_dir, **standard_config)

    def _run_batch(self, batch, fetch='prediction'):
        import torch
        label_dtype = torch.uint8 if fetch == 'prediction_compact' else torch.int64
        probs, variances = [], []
        split = split_samples_over_ranks(self)
        cfgs = {}
        for i, m in enumerate(self.modalities):
            # one call: the dropout-free pass (probabilities, variance_mix.py:68-69) rides along
            # as a leading sample of the MC batch, so conv1_1..pool3 run once for both
            cfgs[m] = {'rate': self.config['dropout_rate'], 'layers': ['pool3'],
                       'num_samples': self.config['num_samples'], 'with_deterministic': True,
                       'seed': mc_dropout_seed(self, i)}
        if split is None:
            outs = self._run_experts(
                batch, lambda m, x: self._experts[self._expert_prefix(m)].forward(
                    x, want=('prob', 'mean_var'), dropout=cfgs[m]), order=self.modalities)
        for i, m in enumerate(self.modalities):
            expert = self._experts[self._expert_prefix(m)]
            cfg = cfgs[m]
            if split is None:
                out = outs[m]
                variances.append(out['mean_var'])
            else:
                # batch-1 latency mode: this rank draws its share of the samples, the per-rank
                # moments are merged over NCCL (every rank ends with the full result)
                cfg.update(num_samples=split[0], seed=cfg['seed'] + 7919 * (split[1] + 1))
                out = expert.forward(batch[m], want=('prob', 'mean_prob', 'var_prob'), dropout=cfg)
                sharding.combine_moments_(out['mean_prob'], out['var_prob'], split[0])
                variances.append(out['var_prob'].mean(-1))
            probs.append(out['prob'])
        score, label = dev.variance_fuse(probs, variances, want_score=(fetch == 'fused_score'),
                                         label_dtype=label_dtype)
        return score if fetch == 'fused_score' else label
