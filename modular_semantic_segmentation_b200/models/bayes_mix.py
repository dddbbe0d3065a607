"""Mirror of xview/models/bayes_mix.py: confusion-matrix Bayes fusion."""
from itertools import product

import numpy as np
import torch

from .. import device as dev
from .basic_fusion_model import FusionModel


def _conditional(confusion_matrix):
    """bayes_mix.py:35: p(expert output | ground-truth class), nan -> 0."""
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.nan_to_num(confusion_matrix / confusion_matrix.sum(0))


def _prior(confusion_matrix, class_prior):
    """bayes_mix.py:42-54, including the constant 1/14 uniform prior and the use of the LAST
    expert's matrix for the data prior."""
    uniform_prior = 1.0 / 14
    with np.errstate(divide='ignore', invalid='ignore'):
        data_prior = confusion_matrix.sum(0) / confusion_matrix.sum()
    if class_prior == 'uniform':
        return uniform_prior
    if class_prior == 'data':
        return data_prior
    weight = float(class_prior)
    prior = weight * uniform_prior + (1 - weight) * data_prior
    return prior / prior.sum()


_TABLE_CACHE = {}


def bayes_tables(confusion_matrices, class_prior='data'):
    """Host-side constants of bayes_fusion (bayes_mix.py:32-54) in the dtype of the matrices:
    log(1e-20 + conditional) per expert [M,C,C] and log(prior) [C]."""
    dtype = np.asarray(confusion_matrices[0]).dtype
    with np.errstate(divide='ignore'):
        log_cond = np.stack([np.log(np.asarray(1e-20, dtype) + _conditional(np.asarray(m)))
                             for m in confusion_matrices])
        log_prior = np.log(np.asarray(_prior(np.asarray(confusion_matrices[-1]), class_prior),
                                      dtype))
    num_classes = log_cond.shape[-1]
    return log_cond, np.broadcast_to(log_prior, (num_classes,)).astype(dtype)


def bayes_fusion(classifications, confusion_matrices, class_prior='data'):
    """bayes_mix.py:12-58 on the device.  classifications: list of int64/uint8 CUDA label maps;
    confusion_matrices: list of numpy arrays (rows = expert output, cols = ground truth).
    Returns (fused score [.., C] float32 CUDA, log-likelihood tables, conditionals)."""
    # the tables are graph constants in the reference; build + upload them once per set of matrices
    key = (tuple(np.asarray(m, np.float32).tobytes() for m in confusion_matrices), str(class_prior))
    if key not in _TABLE_CACHE:
        log_cond, log_prior = bayes_tables([np.asarray(m, np.float32)
                                            for m in confusion_matrices], class_prior)
        conditionals = [_conditional(np.asarray(m, np.float32)) for m in confusion_matrices]
        _TABLE_CACHE[key] = (dev.to_device(log_cond), dev.to_device(log_prior), list(log_cond),
                             conditionals)
    d_cond, d_prior, log_cond, conditionals = _TABLE_CACHE[key]
    score, _ = dev.bayes_fuse_score(classifications, d_cond, d_prior)
    return score, log_cond, conditionals


def bayes_decision_table(confusion_matrices, class_prior='data'):
    """All C^M label combinations evaluated with EXACTLY the arithmetic of `bayes_fusion`
    (same dtype, same summation order: experts in order, then + log prior), so that integer
    lookups on the device reproduce the argmax of the literal per-pixel rule bit for bit."""
    log_cond, log_prior = bayes_tables(confusion_matrices, class_prior)
    num_experts, num_classes = log_cond.shape[0], log_cond.shape[-1]
    combos = np.array(list(product(*(range(num_classes) for _ in range(num_experts)))))
    total = log_cond[0][combos[:, 0]]
    for i in range(1, num_experts):
        total = total + log_cond[i][combos[:, i]]
    total = total + log_prior
    return np.argmax(total, axis=1).reshape([num_classes] * num_experts).astype(np.int32)


def bayes_decision_matrix(confusion_matrices, class_prior='data'):
    """The fused decision for every combination of expert outputs, with the arithmetic of
    bayes_mix.py:61-112: per-expert log(1e-20 + conditional) rows gathered into a FLOAT64
    buffer [combinations, experts, classes] (whatever dtype the matrices have), summed over the
    experts, plus log(prior), argmax.  Returns an array of shape [C] * num_experts."""
    experts = [np.asarray(m) for m in confusion_matrices]
    num_classes, num_experts = experts[0].shape[0], len(experts)
    # row i: expert i's output in every combination, last expert varying fastest
    outputs = np.indices((num_classes,) * num_experts).reshape(num_experts, -1)
    gathered = np.zeros((outputs.shape[1], num_experts, num_classes))
    with np.errstate(divide='ignore'):
        for i, matrix in enumerate(experts):
            gathered[:, i, :] = np.log(1e-20 + _conditional(matrix)[outputs[i]])
        score = gathered.sum(1) + np.log(_prior(experts[-1], class_prior))
    return np.argmax(score, axis=1).reshape((num_classes,) * num_experts)


class BayesFusion(FusionModel):
    """bayes_mix.py:115-161."""

    output_attrs = ('prediction', 'fused_score')
    expert_wants = ('label',)

    def __init__(self, output_dir=None, confusion_matrices=False, **config):
        standard_config = {'learning_rate': 0.0, 'class_prior': 'data'}
        standard_config.update(config)
        self.modalities = []
        self.confusion_matrices = {}
        if confusion_matrices:
            for key, matrix in confusion_matrices.items():
                self.modalities.append(key)
                self.confusion_matrices[key] = np.asarray(matrix).astype('float32').T
        else:
            # bayes_mix.py:143-147 reads the matrices from stored runs (file layout of
            # experiments/utils.py:79-104; folder from config['experiment_storage_folder'] or
            # the XVIEW_EXPERIMENT_STORAGE_FOLDER environment variable).  Unlike the reference,
            # which leaves self.modalities empty on this path, the keys become the modalities.
            from ..records import ExperimentData
            folder = config.get('experiment_storage_folder')
            for key, exp_id in config['eval_experiments'].items():
                self.modalities.append(key)
                matrix = ExperimentData(exp_id, folder).get_record()['info']['confusion_matrix']
                if isinstance(matrix, dict):
                    matrix = matrix['values']
                self.confusion_matrices[key] = np.array(matrix).astype('float32').T
        FusionModel.__init__(self, 'BayesFusion', output_dir=output_dir, **standard_config)

    def _build_graph(self):
        FusionModel._build_graph(self)
        matrices = [self.confusion_matrices[m] for m in self.modalities]
        self.decision_table = bayes_decision_table(matrices, self.config['class_prior'])
        self._lut = dev.to_device(self.decision_table)
        log_cond, log_prior = bayes_tables(matrices, self.config['class_prior'])
        self.likelihoods = list(log_cond)
        self.conditionals = [_conditional(m) for m in matrices]
        self._log_cond = dev.to_device(log_cond)
        self._log_prior = dev.to_device(log_prior)

    def get_insight(self, batch):
        """What experiments/bayes_fusion.py:57-61 (`collect_data`) asks of the model - the
        reference calls this method without defining it; the four entries are the tensors the
        command stores: per-expert softmax probabilities [M,N,H,W,C], the log-likelihood rows
        log(1e-20 + p(output | class)) selected by each expert's decision and the conditionals
        themselves (second and third return value of bayes_fusion, bayes_mix.py:30-58), both
        [M,N,H,W,C], and the fused prediction [N,H,W] - as numpy arrays."""
        batch = self._to_device({k: v for k, v in batch.items() if k != 'labels'})
        outputs = self._expert_outputs(batch, ('prob', 'label'), torch.int64)
        labels = [outputs[m]['classification'] for m in self.modalities]
        fused = dev.bayes_fuse_lut(labels, self._lut, self.config['num_classes'])
        cond = torch.from_numpy(np.stack(self.conditionals).astype(np.float32)).cuda()
        log_cond = self._log_cond.reshape(cond.shape)
        picked = [(log_cond[i][lab], cond[i][lab]) for i, lab in enumerate(labels)]
        return (torch.stack([outputs[m]['prob'] for m in self.modalities]).cpu().numpy(),
                torch.stack([p[0] for p in picked]).cpu().numpy(),
                torch.stack([p[1] for p in picked]).cpu().numpy(),
                fused.cpu().numpy())

    def _fusion(self, expert_outputs, fetch, label_dtype):
        labels = [expert_outputs[m]['classification'] for m in self.modalities]
        if fetch == 'fused_score':
            return dev.bayes_fuse_score(labels, self._log_cond, self._log_prior)[0]
        return dev.bayes_fuse_lut(labels, self._lut, self.config['num_classes'])

    def _score_batch(self, batch, cm):
        """score() step with the tail fused into one kernel: each expert runs up to its
        low-resolution class scores, then label decode + decision table + confusion matrix
        happen in a single pass (no label map in HBM).  Needs the bilinear decoder fast path of
        the FCN experts; anything else takes the generic route."""
        if self.config.get('fused_score_tail', True) and not getattr(self, '_no_fused_tail', False) \
                and self.config['expert_model'] == 'fcn' and len(self.modalities) <= 4 \
                and self.config.get('precision', 'bf16') == 'bf16':
            experts = [self._experts[self._expert_prefix(m)] for m in self.modalities]
            self._run_experts(batch, lambda m, x: self._experts[self._expert_prefix(m)].forward(
                x, want=()))
            try:
                dev.bayes_decode_score(experts, self._lut, self.config['num_classes'],
                                       batch['labels'], cm)
                return
            except dev._abi.XViewError:
                self._no_fused_tail = True      # e.g. non-bilinear upscore kernels were imported
        FusionModel._score_batch(self, batch, cm)
