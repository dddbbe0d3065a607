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

    def make_trainer(self):
        """The training step of this model as an object (fit() drives it; bench.py times it)."""
        return Trainer(self)

    def fit(self, dataset, iterations, output=True, validation_dataset=None,
            validation_interval=100, additional_eval_datasets={}):
        """base_model.py:179-261: `iterations` steps of config['trainer'] ('adam' | 'adagrad' |
        'rmsprop', TensorFlow 1.x defaults, base_model.py:157-162) at config['learning_rate'] on
        the cross-entropy of simple_fcn.py:205-215.  Every `validation_interval` steps the
        validation set is scored and `loss`, `accuracy`, `IoU` (plus the mean IoU of every
        `additional_eval_datasets` entry) are appended to `<output_dir>/summaries.jsonl` - the
        TensorBoard summaries of base_model.py:191-251 without TensorFlow.  Batches are uploaded
        one step ahead on a copy stream.  With torch.distributed initialised every rank trains
        on its share of each batch; the gradient buckets are all-reduced over NCCL while the
        backward pass is still running (Trainer.step).  With batch_normalization=True every layer
        normalises with the statistics of the rank's own batch (as the reference's replicas
        would: no synchronised batch norm) and the moving statistics are updated every step
        (the UPDATE_OPS dependency of base_model.py:155-156)."""
        import json
        import os
        trainer = self.make_trainer()
        batches = self._device_batches(self._training_batches(dataset), presharded=True)
        summaries = None
        if self.output_dir is not None:
            os.makedirs(self.output_dir, exist_ok=True)
            summaries = open(os.path.join(self.output_dir, 'summaries.jsonl'), 'a')
        print('INFO: Start training')
        self.loss_history = []
        try:
            for i in range(iterations):
                trainer.step(next(batches))
                self.global_step += 1
                if validation_dataset is not None and i % validation_interval == 0:
                    record = {'step': i, 'loss': trainer.last_loss()}
                    self.loss_history.append(record['loss'])
                    trainer.sync_variables()
                    score, _ = self.score(validation_dataset)
                    record['accuracy'] = float(score['total_accuracy'])
                    record['IoU'] = float(score['mean_IoU'])
                    for key, extra in additional_eval_datasets.items():
                        record[key] = float(self.score(extra)[0]['mean_IoU'])
                    if output:
                        print("{:4d}: accuracy {:.2f}, IoU {:.2f}".format(
                            i, score['total_accuracy'], score['mean_IoU']))
                    if summaries is not None:
                        summaries.write(json.dumps(record) + '\n')
                        summaries.flush()
                    if 'abort_at_iou' in self.config and \
                            score['mean_IoU'] > self.config['abort_at_iou']:
                        break
        finally:
            if summaries is not None:
                summaries.close()
            batches.close()
        if iterations > 0:
            self.loss = trainer.last_loss()
        trainer.close()
        print('INFO: Training finished.')

    def _pull_trained_variables(self):
        """Device master parameters -> self.variables (what export_weights writes) and back into
        the expert's host-side parameter store, so that a later finalize() / fit() continues
        from the trained weights (the optimizer state on the device is kept)."""
        expert = self._experts[self.prefix]
        flat = expert.get_params()
        head = self.prefix + '/'
        updated = {}
        for name in list(self.variables.keys()):
            below = name.split('/', 1)[1]
            try:
                off, size = expert.param_span(below)
            except Exception:
                continue                               # non-trainable (bilinear kernels)
            self.variables[name] = flat[off:off + size].reshape(self.variables[name].shape).copy()
            updated[name[len(head):]] = self.variables[name]
        expert.set_params(updated, keep_train_state=True)

    def _run_batch(self, batch, fetch='prediction'):
        expert = self._experts[self.prefix]
        x = batch[self.modality]
        if fetch == 'prediction':
            return expert.forward(x, want=('label',))['label']
        if fetch == 'prediction_compact':
            return expert.forward(x, want=('label',), label_dtype=torch.uint8)['label']
        return expert.forward(x, want=(fetch,))[fetch]


class Trainer(object):
    """One optimisation step of SimpleFCN.fit (base_model.py:153-162 + simple_fcn.py:205-215).

    step(batch): forward + backward on this rank's batch (dict of CUDA tensors), sum of the
    un-normalised gradients and of {loss sum, valid-pixel count} over the ranks, division by the
    global pixel count, optimizer step.  Data parallel: the flat gradient is cut into the
    buckets of `FcnExpert.grad_buckets()`; the backward pass records an event as each bucket
    becomes final and the bucket's NCCL all-reduce is issued on a communication stream right
    away, so that only the last (smallest) bucket's reduction is exposed."""

    def __init__(self, model):
        from .. import sharding
        self.model = model
        self.expert = model._experts[model.prefix]
        self.trainer = model.config.get('trainer', 'adam')
        if self.trainer not in dev.FcnExpert.OPTIMIZERS:
            raise UserWarning('ERROR: unknown trainer %r' % (self.trainer,))
        self.learning_rate = float(model.config.get('learning_rate', 0.0001))
        self.train_encoder = bool(model.config.get('train_encoder', True))
        if not self.expert._train_live:        # a second fit() continues: state is kept
            self.expert.train_begin()
        self.dist = sharding.dist_or_none()
        self.grads = self.loss = None
        self.buckets = self.events = self.comm = None
        if self.dist is not None and model.config.get('overlap_allreduce', True):
            self.buckets = self.expert.grad_buckets()
            self.events = [torch.cuda.Event() for _ in self.buckets]
            self.comm = torch.cuda.Stream()

    def step(self, batch):
        from .. import sharding
        model = self.model
        x, labels = batch[model.modality], batch['labels']
        self.grads, self.loss = self.expert.train_gradients(
            x, labels, train_encoder=self.train_encoder, normalize=False, grads=self.grads,
            loss=self.loss, bucket_events=self.events)
        if self.dist is not None:
            compute = torch.cuda.current_stream()
            if self.events is None:
                sharding.allreduce_sum_(self.grads)          # one flat bucket, not overlapped
                sharding.allreduce_sum_(self.loss)
            else:
                with torch.cuda.stream(self.comm):
                    for (start, stop), event in zip(self.buckets, self.events):
                        self.comm.wait_event(event)
                        self.dist.all_reduce(self.grads[start:stop])
                    self.comm.wait_stream(compute)           # the loss is written last
                    self.dist.all_reduce(self.loss)
                compute.wait_stream(self.comm)
        dev.scale_by_count(self.grads, self.loss)
        self.expert.optimizer_step(self.grads, self.trainer, self.learning_rate)

    def last_loss(self):
        """Mean cross-entropy over the valid pixels of the last step (all ranks)."""
        l = self.loss.cpu().numpy()
        return float(l[0] / (1e-20 + l[1]))

    def sync_variables(self):
        self.model._pull_trained_variables()

    def close(self):
        self.sync_variables()
