"""Mirror of xview/models/simple_fcn.py: the VGG16-FCN expert, functional `fcn()` and the
`SimpleFCN` model class, running on the B200 kernels."""
from collections import OrderedDict

import numpy as np
import torch

from .. import device as dev
from .base_model import BaseModel
from .custom_layers import bilinear_filter_initializer, glorot_uniform

CONV_LAYERS = [('conv1_1', 64), ('conv1_2', 64), ('conv2_1', 128), ('conv2_2', 128),
               ('conv3_1', 256), ('conv3_2', 256), ('conv3_3', 256), ('conv4_1', 512),
               ('conv4_2', 512), ('conv4_3', 512), ('conv5_1', 512), ('conv5_2', 512),
               ('conv5_3', 512)]


def init_fcn_variables(prefix, num_channels, num_units, num_classes, batchnorm=False, rng=None):
    """Variables of one expert under their TensorFlow names with the initial values
    `tf.global_variables_initializer()` gives them in the reference: Glorot-uniform kernels,
    zero biases, bilinear transposed-conv kernels, identity batch-norm statistics
    (simple_fcn.py:39-83,129-133; layout: SURVEY.md Appendix B)."""
    rng = rng if rng is not None else np.random.default_rng()
    v = OrderedDict()

    def bn(scope, c):
        if batchnorm:
            v[scope + '/gamma'] = np.ones(c, np.float32)
            v[scope + '/beta'] = np.zeros(c, np.float32)
            v[scope + '/moving_mean'] = np.zeros(c, np.float32)
            v[scope + '/moving_variance'] = np.ones(c, np.float32)

    cin = num_channels
    for name, cout in CONV_LAYERS:
        scope = '%s/%s' % (prefix, name)
        v[scope + '/kernel'] = glorot_uniform((3, 3, cin, cout), rng)
        v[scope + '/bias'] = np.zeros(cout, np.float32)
        bn(scope, cout)
        cin = cout
    for name in ('score_conv4', 'score_conv5'):
        scope = '%s/%s' % (prefix, name)
        v[scope + '/kernel'] = glorot_uniform((1, 1, 512, num_units), rng)
        v[scope + '/bias'] = np.zeros(num_units, np.float32)
        bn(scope, num_units)
    v[prefix + '/upscore_conv5/kernel'] = bilinear_filter_initializer(
        (4, 4, num_units, num_units))
    bn(prefix + '/upscore_conv5', num_units)
    v[prefix + '/upscore/kernel'] = bilinear_filter_initializer((16, 16, num_units, num_units))
    bn(prefix + '/upscore', num_units)
    v[prefix + '/score/kernel'] = glorot_uniform((1, 1, num_units, num_classes), rng)
    v[prefix + '/score/bias'] = np.zeros(num_classes, np.float32)
    bn(prefix + '/score', num_classes)
    return v


def build_expert(prefix, num_channels, num_units, num_classes, batchnorm=False,
                 precision='bf16', rng=None):
    """Creates the device expert plus its freshly initialised variables."""
    variables = init_fcn_variables(prefix, num_channels, num_units, num_classes, batchnorm, rng)
    expert = dev.FcnExpert(num_channels, num_units, num_classes, batchnorm=batchnorm,
                           precision=precision)
    return expert, variables


# variable store of the functional API: the TF variable scopes of simple_fcn.py:36 with
# reuse=tf.AUTO_REUSE become (prefix -> expert) entries
_VARIABLE_STORE = {}


def reset_variable_store():
    for expert, _ in _VARIABLE_STORE.values():
        expert.close()
    _VARIABLE_STORE.clear()


def fcn(inputs, prefix, num_units, num_classes, trainable=True, is_training=False, reuse=None,
        dropout_rate=0, dropout_layers=[], batchnorm=True, num_samples=1, seed=0, masks=None,
        precision='bf16', want=('score',), params=None, label_dtype=torch.int64):
    """simple_fcn.py:137-170 as called by experiments/timing.py:55-58 and the fusion models.

    inputs: CUDA float32 [N,H,W,Cin].  Returns a dict with the requested entries of
    'score', 'prob', 'label' (= 'classification'), and for num_samples > 1 'mean_prob',
    'var_prob', 'mean_var'.  Variables are created on first use of `prefix` (random init as in
    the reference) and reused afterwards; `params` overrides them (names below the prefix)."""
    key = (prefix, int(inputs.shape[-1]), num_units, num_classes, bool(batchnorm), precision)
    if key not in _VARIABLE_STORE:
        expert, variables = build_expert(prefix, key[1], num_units, num_classes, batchnorm,
                                         precision)
        expert.set_params({n.split('/', 1)[1]: a for n, a in variables.items()})
        _VARIABLE_STORE[key] = (expert, variables)
    expert, variables = _VARIABLE_STORE[key]
    if params is not None:
        expert.set_params(params)
    dropout = None
    if dropout_layers and dropout_rate:
        dropout = {'rate': dropout_rate, 'layers': list(dropout_layers),
                   'num_samples': num_samples, 'seed': seed, 'masks': masks}
    out = expert.forward(inputs, want=want, dropout=dropout, label_dtype=label_dtype)
    if 'label' in out:
        out['classification'] = out['label']
    return out


class SimpleFCN(BaseModel):
    """simple_fcn.py:173-224.  Test-time network: softmax + argmax of the FCN scores."""

    output_attrs = ('prediction', 'prob', 'score')

    def __init__(self, prefix, data_description, modality, output_dir=None, **config):
        self.prefix = prefix
        self.modality = modality
        standard_config = {'train_encoder': True, 'dropout_rate': 0}
        standard_config.update(config)
        BaseModel.__init__(self, data_description, output_dir=output_dir, **standard_config)

    def _build_graph(self):
        channels = self.testdata_description[1][self.modality][-1]
        expert, variables = build_expert(
            self.prefix, channels, self.config['num_units'], self.config['num_classes'],
            batchnorm=self.config['batch_normalization'],
            precision=self.config.get('precision', 'bf16'),
            rng=np.random.default_rng(self.config.get('seed')))
        self._register_expert(self.prefix, expert, variables)
        self.prediction = 'prediction'

    # ------------------------------------------------------------------ training
    def _training_batches(self, dataset):
        """Endless stream of training batches (the `.repeat().batch()` of base_model.py:203-206):
        a dict of arrays is cycled in order, any other iterable of batch dicts is restarted when
        it is exhausted (pass a list or a re-iterable object)."""
        while True:
            empty = True
            for batch in self._batches(dataset):
                empty = False
                yield batch
            if empty:
                raise UserWarning('ERROR: empty training dataset')

    def fit(self, dataset, iterations, output=True, validation_dataset=None,
            validation_interval=100, additional_eval_datasets={}):
        """base_model.py:179-261: `iterations` Adam steps (trainer / learning_rate from the
        config, tf defaults beta1=0.9, beta2=0.999, eps=1e-8) on the cross-entropy of
        simple_fcn.py:205-215.  With torch.distributed initialised every rank trains on its
        share of each batch and the gradients are summed over ranks (one bucketed all-reduce
        of the flat gradient vector) before the shared Adam step."""
        from .. import sharding
        if self.config.get('trainer', 'adam') != 'adam':
            raise UserWarning('ERROR: only the adam trainer is available on the B200 path')
        if self.config['batch_normalization']:
            raise UserWarning('ERROR: fit() with batch normalisation is not built yet')
        expert = self._experts[self.prefix]
        expert.train_begin()
        lr = float(self.config.get('learning_rate', 0.0001))
        train_encoder = bool(self.config.get('train_encoder', True))
        batches = self._training_batches(dataset)
        grads = loss = None
        print('INFO: Start training')
        self.loss_history = []
        for i in range(iterations):
            batch = self._to_device(next(batches))
            grads, loss = expert.train_gradients(batch[self.modality], batch['labels'],
                                                 train_encoder=train_encoder, normalize=False,
                                                 grads=grads, loss=loss)
            sharding.allreduce_sum_(grads)           # bucketed: one flat tensor
            sharding.allreduce_sum_(loss)
            dev.scale_by_count(grads, loss)
            expert.adam_step(grads, learning_rate=lr)
            self.global_step += 1
            if validation_dataset is not None and i % validation_interval == 0:
                l = loss.cpu().numpy()
                self.loss_history.append(float(l[0] / (1e-20 + l[1])))
                self._pull_trained_variables()
                score, _ = self.score(validation_dataset)
                if output:
                    print("{:4d}: accuracy {:.2f}, IoU {:.2f}".format(
                        i, score['total_accuracy'], score['mean_IoU']))
                if 'abort_at_iou' in self.config and \
                        score['mean_IoU'] > self.config['abort_at_iou']:
                    break
        if loss is not None:
            l = loss.cpu().numpy()
            self.loss = float(l[0] / (1e-20 + l[1]))
        self._pull_trained_variables()
        print('INFO: Training finished.')

    def _pull_trained_variables(self):
        """Device master parameters -> self.variables (what export_weights writes)."""
        expert = self._experts[self.prefix]
        flat = expert.get_params()
        for name in list(self.variables.keys()):
            below = name.split('/', 1)[1]
            try:
                off, size = expert.param_span(below)
            except Exception:
                continue                               # non-trainable (bilinear kernels)
            self.variables[name] = flat[off:off + size].reshape(self.variables[name].shape).copy()

    def _run_batch(self, batch, fetch='prediction'):
        expert = self._experts[self.prefix]
        x = batch[self.modality]
        if fetch == 'prediction':
            return expert.forward(x, want=('label',))['label']
        if fetch == 'prediction_compact':
            return expert.forward(x, want=('label',), label_dtype=torch.uint8)['label']
        return expert.forward(x, want=(fetch,))[fetch]
