"""Mirror of xview/models/dirichlet_mix.py: Dirichlet fusion of the experts' softmax outputs."""
from copy import deepcopy

import numpy as np
import torch
from scipy.special import gammaln

from .. import device as dev
from .. import sharding
from .base_model import BaseModel
from .basic_fusion_model import build_test_pipeline
from .dirichletDifferentiation import findDirichletPriors


def dirichlet_tables(dirichlet_params, sigma, prior):
    """Constants of dirichlet_fusion (dirichlet_mix.py:107-113,34-36) for the device kernel:
    alpha-1 [M,C_out,C_gt], lbeta(alpha[:,c]) [M,C_gt] and log(1e-20+prior) [C], float32.
    Column c of a parameter matrix is the concentration for ground-truth class c."""
    alpha = [np.float32(sigma) * np.asarray(p).astype('float32') for p in dirichlet_params]
    alpha_m1 = np.stack([a - np.float32(1) for a in alpha]).astype(np.float32)
    log_norm = np.stack([(gammaln(a.astype(np.float64)).sum(0) -
                          gammaln(a.astype(np.float64).sum(0))) for a in alpha]).astype(np.float32)
    num_classes = alpha_m1.shape[-1]
    prior = np.broadcast_to(np.asarray(prior, np.float32), (num_classes,))
    with np.errstate(divide='ignore'):
        log_prior = np.log(np.float32(1e-20) + prior).astype(np.float32)
    return alpha_m1, log_norm, log_prior


def class_prior_from_counts(class_counts, class_prior):
    """dirichlet_mix.py:116-129 (uniform prior is the constant 1/14)."""
    class_counts = np.asarray(class_counts).astype('float32')
    uniform_prior = 1.0 / 14
    data_prior = (class_counts / (1e-20 + class_counts.sum())).astype('float32')
    if class_prior == 'uniform':
        return uniform_prior
    if class_prior == 'data':
        return data_prior
    weight = float(class_prior)
    prior = weight * uniform_prior + (1 - weight) * data_prior
    return prior / prior.sum()


def dirichlet_fusion(probs, dirichlet_params, prior, sigma=1.0, want_score=True, exact=True):
    """dirichlet_mix.py:14-36 on the device.  probs: list of CUDA float32 [N,H,W,C] softmax
    outputs; dirichlet_params: list of [C,C] numpy arrays; prior: [C] or scalar.  Returns the
    fused score [N,H,W,C] (argmax over the last axis is the fused classification).
    exact=True (default): the fixed-order float32 arithmetic whose argmax is bit-exact."""
    alpha_m1, log_norm, log_prior = dirichlet_tables(dirichlet_params, sigma, prior)
    score, label = dev.dirichlet_fuse(
        probs, dev.to_device(alpha_m1), dev.to_device(log_norm), dev.to_device(log_prior),
        want_score=want_score, exact=exact,
        magnitudes=dev.dirichlet_table_magnitudes(alpha_m1, log_norm, log_prior))
    return score if want_score else label


class DirichletFusion(BaseModel):
    """dirichlet_mix.py:39-273.  The modality name doubles as the variable prefix (:98)."""

    output_attrs = ('prediction', 'fused_score')

    def __init__(self, output_dir=None, **config):
        standard_config = {'learning_rate': 0.0, 'sigma': 1.0, 'class_prior': 'data',
                           'delta': 1e-2, 'beta': 1e-2}
        standard_config.update(config)
        self.modalities = config['modalities']
        if 'measurement_exp' in config or 'dirichlet_params' in config:
            if 'measurement_exp' in config:
                # dirichlet_mix.py:62-64: the fit stored by a measurement run (counts.npz)
                import io
                from ..records import ExperimentData
                stored = ExperimentData(config['measurement_exp'],
                                        config.get('experiment_storage_folder'))
                with stored.get_artifact('counts.npz') as f:
                    measurements = dict(np.load(io.BytesIO(f.read())))
            else:
                measurements = config['dirichlet_params']
            self.dirichlet_params = {m: np.asarray(measurements[m]).astype('float32')
                                     for m in self.modalities}
            self.class_counts = np.asarray(measurements['class_counts']).astype('float32')
        else:
            print('WARNING: Could not yet import measurements, you need to fit this '
                  'model first.')
        BaseModel.__init__(self, name='DirichletFusion', output_dir=output_dir,
                           custom_training=True, **standard_config)

    def _build_graph(self):
        if not self._experts:       # re-entered by fit(): keep the experts and their weights
            for m in self.modalities:
                expert, variables = build_test_pipeline(m, self.config['num_channels'][m],
                                                        **self.config)
                self._register_expert(m, expert, variables)
        if hasattr(self, 'dirichlet_params'):
            prior = class_prior_from_counts(self.class_counts, self.config['class_prior'])
            tables = dirichlet_tables([self.dirichlet_params[m] for m in self.modalities],
                                      self.config['sigma'], prior)
            self._tables = [dev.to_device(t) for t in tables]
            self._magnitudes = dev.dirichlet_table_magnitudes(*tables)
            self.prediction = 'prediction'
        else:
            self._tables = None
            self.prediction = 0     # dirichlet_mix.py:166-167: nothing to fuse before fit()

    def _probs(self, batch):
        """Expert outputs that enter the fusion.  `num_samples` > 1 (BASELINE configs[2]):
        the mean softmax of that many MC-dropout passes per modality (dropout at
        `dropout_layers`, default after pool3 as variance_mix.py:56), all samples sharing one
        weight load; otherwise the dropout-free softmax of dirichlet_mix.py:98."""
        num_samples = int(self.config.get('num_samples', 1))
        if num_samples <= 1:
            outs = self._run_experts(
                batch, lambda m, x: self._experts[m].forward(x, want=('prob',))['prob'])
            return [outs[m] for m in self.modalities]
        from .variance_mix import mc_dropout_seed, split_samples_over_ranks
        split = split_samples_over_ranks(self)
        means = []
        for i, m in enumerate(self.modalities):
            cfg = {'rate': self.config.get('dropout_rate', 0.5),
                   'layers': list(self.config.get('dropout_layers', ['pool3'])),
                   'num_samples': num_samples, 'seed': mc_dropout_seed(self, i)}
            if split is None:
                means.append(self._experts[m].forward(batch[m], want=('mean_prob',),
                                                      dropout=cfg)['mean_prob'])
            else:       # samples split over the ranks (batch-1 latency mode), moments merged
                cfg.update(num_samples=split[0], seed=cfg['seed'] + 7919 * (split[1] + 1))
                out = self._experts[m].forward(batch[m], want=('mean_prob', 'var_prob'),
                                               dropout=cfg)
                sharding.combine_moments_(out['mean_prob'], out['var_prob'], split[0])
                means.append(out['mean_prob'])
        return means

    def _fused_tail_applies(self):
        """The one-kernel tail (decode of both experts + Dirichlet fusion [+ confusion matrix],
        xv_dirichlet_decode_score) covers the reference's own configuration: two FCN experts,
        dropout-free softmax outputs, up to 16 classes."""
        return (self.config.get('fused_score_tail', True) and
                not getattr(self, '_no_fused_tail', False) and
                self.config['expert_model'] == 'fcn' and len(self.modalities) == 2 and
                self.config.get('precision', 'bf16') == 'bf16' and
                int(self.config.get('num_samples', 1)) <= 1 and self.config['num_classes'] <= 16)

    def _fused_tail(self, batch, cm, want_label, label_dtype=torch.int64):
        experts = [self._experts[m] for m in self.modalities]
        self._run_experts(batch, lambda m, x: self._experts[m].forward(x, want=()))
        try:
            return True, dev.dirichlet_decode_score(
                experts, *self._tables, self.config['num_classes'], self._magnitudes,
                gt_labels=batch['labels'].contiguous() if cm is not None else None, cm=cm,
                want_label=want_label, label_dtype=label_dtype,
                exact=bool(self.config.get('exact_fusion', True)))
        except dev._abi.XViewError:
            self._no_fused_tail = True      # e.g. non-bilinear upscore kernels were imported
            return False, None

    def _score_batch(self, batch, cm):
        if self._tables is not None and self._fused_tail_applies():
            done, _ = self._fused_tail(batch, cm, want_label=False)
            if done:
                return
        BaseModel._score_batch(self, batch, cm)

    def _run_batch(self, batch, fetch='prediction'):
        if self._tables is None:
            raise UserWarning('ERROR: DirichletFusion has to be fitted before inference')
        label_dtype = torch.uint8 if fetch == 'prediction_compact' else torch.int64
        if fetch in ('prediction', 'prediction_compact') and self._fused_tail_applies():
            done, label = self._fused_tail(batch, None, want_label=True, label_dtype=label_dtype)
            if done:
                return label
        self.probs = dict(zip(self.modalities, self._probs(batch)))
        # `exact_fusion` (default True): bit-exact argmax of the fixed-order float32 rule; the
        # fast arithmetic is kept for pixels whose decision margin exceeds the error bound
        score, label = dev.dirichlet_fuse([self.probs[m] for m in self.modalities],
                                          *self._tables, want_score=(fetch == 'fused_score'),
                                          label_dtype=label_dtype,
                                          exact=bool(self.config.get('exact_fusion', True)),
                                          magnitudes=self._magnitudes)
        return score if fetch == 'fused_score' else label

    def _get_sufficient_statistic(self, data):
        """dirichlet_mix.py:173-205: per modality S[c,k] = sum_{label==c} log(1e-10+prob[k])
        and the class counts, accumulated on the device (float64 / int64) over all batches and
        summed over ranks."""
        c = self.config['num_classes']
        stats = {m: torch.zeros((c, c), dtype=torch.float64, device='cuda')
                 for m in self.modalities}
        counts = torch.zeros(c, dtype=torch.int64, device='cuda')
        scratch = torch.zeros(c, dtype=torch.int64, device='cuda')
        for batch in self._device_batches(data):
            labels = batch['labels'].contiguous()
            for i, (m, prob) in enumerate(zip(self.modalities, self._probs(batch))):
                dev.dirichlet_suffstats(prob, labels, stats[m], counts if i == 0 else scratch)
        if not self._same_images_on_every_rank():
            for m in self.modalities:
                sharding.allreduce_sum_(stats[m])
            sharding.allreduce_sum_(counts)
        return {m: s.cpu().numpy() for m, s in stats.items()}, counts.cpu().numpy()

    def _fit_sufficient_statistic(self, counts, class_counts):
        """dirichlet_mix.py:207-257 (host, float64)."""
        num_classes = self.config['num_classes']

        total_count = class_counts.sum()

        def fit_columns(log_sums):
            """Column c = concentration parameters of the expert's output given ground-truth
            class c: maximum of the regularised likelihood of dirichletDifferentiation.py with
            the mean log-probability inside the class as the statistic and the mean over all
            other classes as the negative statistic.  Classes never seen keep alpha = 1."""
            alpha = np.ones((num_classes, num_classes), np.float64)
            everything = log_sums.sum(0)
            for c in np.flatnonzero(np.asarray(class_counts) != 0):
                inside = np.asarray(log_sums[c, :] / class_counts[c], np.float64)
                outside = (everything - log_sums[c, :]) / (total_count - class_counts[c])
                alpha[:, c] = findDirichletPriors(inside, outside, np.ones(num_classes, np.float64),
                                                  max_iter=10000, delta=self.config['delta'],
                                                  beta=self.config['beta'])
            return alpha

        self.dirichlet_params = {m: fit_columns(counts[m]) for m in self.modalities}
        self.class_counts = class_counts
        self._build_graph()

    def fit(self, data, *args, **kwargs):
        """dirichlet_mix.py:259-273: measure the experts on `data`, fit the class-conditional
        Dirichlets, return {modality: params [C,C], 'class_counts': [C]}."""
        modality_counts, class_counts = self._get_sufficient_statistic(data)
        print('INFO: Measurements of classifiers finished, now EM')
        self._fit_sufficient_statistic(modality_counts, class_counts)
        print("INFO: MixFCN fitted to data")
        return_dict = deepcopy(self.dirichlet_params)
        return_dict['class_counts'] = self.class_counts
        return return_dict
