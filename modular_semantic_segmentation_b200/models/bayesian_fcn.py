"""Mirror of xview/models/bayesian_fcn.py: MC-dropout sampling of an FCN expert and the
uncertainty measures derived from the samples.  The reference class derives from an
`UncertaintyModel` that is not shipped (bayesian_fcn.py:3), so the
working statement is the function `sampling_uncertainty` and the test-time half of
`BayesianFCN._build_graph` (bayesian_fcn.py:9-57, 106-111); the training half is SimpleFCN's."""
import torch

from .. import device as dev
from .simple_fcn import SimpleFCN

DEFAULT_DROPOUT_LAYERS = ('pool3', 'pool4', 'conv4_3', 'conv5_3', 'features')


def _sites(layers):
    # pool4's dropout is gated by 'pool3' in the reference graph (simple_fcn.py:61)
    return [name for name in layers if name != 'pool4']


def sampling_uncertainty(inputs, expert, num_samples, num_classes=None, dropout_rate=0.5,
                         dropout_layers=DEFAULT_DROPOUT_LAYERS, seed=0):
    """bayesian_fcn.py:9-57 on the device: `num_samples` MC-dropout passes of `expert` (an
    FcnExpert handle; all passes share one weight load and the trunk in front of the first
    dropout site) -> (mean probability [N,H,W,C], {'entropy': entropy of the mean,
    'cond_entropy': mean entropy of the samples, 'variance': sum over classes of the population
    variance}), float32 CUDA tensors.  `num_classes` is accepted for signature parity."""
    samples = expert.forward(inputs, want=('prob',),
                             dropout={'rate': dropout_rate, 'layers': _sites(dropout_layers),
                                      'num_samples': int(num_samples), 'seed': seed})['prob']
    # the expert returns the samples stacked on the batch axis, sample-major: [T*N,H,W,C]
    samples = samples.reshape((int(num_samples), -1) + tuple(samples.shape[1:]))
    stats = dev.mc_moments(samples, want=('mean', 'entropy', 'cond_entropy', 'sum_var'))
    return stats['mean'], {'entropy': stats['entropy'], 'cond_entropy': stats['cond_entropy'],
                           'variance': stats['sum_var']}


class BayesianFCN(SimpleFCN):
    """bayesian_fcn.py:60-111, method 'sampling': prediction = argmax of the MC mean; `entropy`,
    `cond_entropy` and `variance` are further outputs of predict(output_attr=...)."""

    output_attrs = ('prediction', 'mean_prob', 'entropy', 'cond_entropy', 'variance')

    def __init__(self, prefix, data_description, modality, output_dir=None,
                 dropout_layers=DEFAULT_DROPOUT_LAYERS, **config):
        standard_config = {'num_samples': 20, 'dropout_rate': 0.5, 'method': 'sampling',
                           'batch_normalization': False}
        standard_config.update(config)
        if standard_config['method'] != 'sampling':
            raise UserWarning('ERROR: only the sampling method is defined (bayesian_fcn.py:89)')
        SimpleFCN.__init__(self, prefix, data_description, modality, output_dir=output_dir,
                           dropout_layers=list(dropout_layers), **standard_config)

    def _run_batch(self, batch, fetch='prediction'):
        from .variance_mix import mc_dropout_seed
        mean, uncertainties = sampling_uncertainty(
            batch[self.modality], self._experts[self.prefix], self.config['num_samples'],
            self.config['num_classes'], dropout_rate=self.config['dropout_rate'],
            dropout_layers=self.config['dropout_layers'], seed=mc_dropout_seed(self, 0))
        if fetch in uncertainties:
            return uncertainties[fetch]
        if fetch == 'mean_prob':
            return mean
        # argmax with the first-maximum tie rule of tf.argmax (the average-fusion kernel over one map)
        label_dtype = torch.uint8 if fetch == 'prediction_compact' else torch.int64
        return dev.average_fuse([mean], label_dtype=label_dtype)[1]
