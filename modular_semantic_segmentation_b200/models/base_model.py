"""Host-side mirror of xview/models/base_model.py (the sklearn-style model surface).

Same constructor, same methods and return types as the reference `BaseModel`
(base_model.py:51-451): `with Model(data_description, **config) as net`, `fit`, `predict`,
`score`, `import_weights`, `export_weights`, `load_weights`, `close`.  The TensorFlow graph +
session are replaced by device-resident experts driven through the C ABI
(`modular_semantic_segmentation_b200.device`); variables are kept as host numpy arrays under
their TensorFlow names so the npz layout of the reference round-trips.

Multi-GPU: if `torch.distributed` is initialised (one process per GPU), `score()` and
`predict()` shard the images over the ranks; `score()` all-reduces the int64 confusion matrix
once at the end (the only collective on the path), `predict()` all-gathers the label maps.
"""
from collections import OrderedDict
from copy import deepcopy
from os import path

import numpy as np
import torch

from .. import device as dev
from .. import sharding
from ..input_pipeline import crop_multiple
from ..sharding import dist_or_none as _dist


def measures_from_confusion_matrix(confusion_matrix):
    """The measures score() reports, in float64 and with the operation order of
    base_model.py:315-329 (they are pinned bit for bit by tests/golden/exp868.npz): rows are
    ground truth, columns predictions; class 0 (void) counts for neither total_accuracy nor
    mean_IoU; classes that never occur give nan and are skipped by the means."""
    cm = np.asarray(confusion_matrix, dtype=np.float64)
    hits = np.diag(cm)
    per_truth, per_prediction = cm.sum(1), cm.sum(0)
    with np.errstate(divide='ignore', invalid='ignore'):
        recall = hits / per_truth
        precision = hits / per_prediction
        f1 = 2 * precision * recall / (precision + recall)
        iou = hits / (per_truth + per_prediction - hits)
        accuracy = hits[1:].sum() / cm[1:, :].sum()
    return {'confusion_matrix': cm, 'recall': recall, 'precision': precision, 'F1': f1,
            'mean_F1': np.nanmean(f1), 'total_accuracy': accuracy, 'IoU': iou,
            'mean_IoU': np.nanmean(iou[1:])}


def upload_bounds(count, split=2, explicit=None):
    """Where a host batch of `count` images is cut for uploading, or None for one upload.
    Every piece costs ~0.5 ms of launch / tail overhead (measured: a fused step takes about
    0.5 + 0.26 n ms for n frames), so there are few pieces; the first one - whose copy nothing
    can hide - is the smallest: [n/4, 3n/4] for split = 2.  `explicit` gives the piece sizes."""
    if explicit and sum(explicit) == count:
        bounds = [0]
        for size in explicit:
            bounds.append(bounds[-1] + int(size))
        return bounds
    if split > 1 and count >= 8 * split:
        lead = max(1, count // (2 * split))
        rest = count - lead
        return [0] + [lead + (rest * i + split - 2) // (split - 1) for i in range(split)]
    return None


def _nbytes(value):
    if isinstance(value, torch.Tensor):
        return value.numel() * value.element_size()
    return getattr(value, 'nbytes', 0)


def upload_order(batch):
    """Keys of a host batch in the order they are copied to the device: image arrays by
    ascending size, labels last (they are needed last)."""
    return sorted(batch, key=lambda k: (k == 'labels', _nbytes(batch[k])))


class _DeviceBatch(dict):
    """Batch of CUDA tensors whose uploads may still be in flight on the copy stream: reading an
    entry makes the compute stream wait for that entry's copy only."""

    def __init__(self, tensors, events, stream):
        dict.__init__(self, tensors)
        self._events = dict(events)
        self._stream = stream

    def _arrived(self, key=None):
        for k in ([key] if key is not None else list(self._events)):
            event = self._events.pop(k, None)
            if event is not None:
                self._stream.wait_event(event)

    def __getitem__(self, key):
        self._arrived(key)
        return dict.__getitem__(self, key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        self._arrived()
        return dict.items(self)

    def values(self):
        self._arrived()
        return dict.values(self)


class BaseModel(object):
    """Structure for network models.  Subclasses implement `_build_graph()` (create the
    experts and register their variables) and `_run_batch(batch, fetch)`."""

    required_attributes = [["loss"], ["prediction"]]

    def __init__(self, data_description, name=None, output_dir=None, custom_training=False,
                 batchsize=1, **config):
        self.name = type(self).__name__ if name is None else name
        self.output_dir = output_dir
        self.custom_training = custom_training
        self.config = config
        self.config['batchsize'] = batchsize
        self.config['num_classes'] = data_description[2]
        # same bookkeeping as base_model.py:84-94
        self.testdata_description = [data_description[0], {
            key: [None, *description] for key, description in data_description[1].items()}]
        train_description = deepcopy(self.testdata_description[1])
        if 'labels' in train_description:
            train_description['labels'] = list(train_description['labels'])
            train_description['labels'].append(self.config['num_classes'])
        self.data_description = [self.testdata_description[0], train_description]
        self.variables = OrderedDict()     # TF variable name -> float32 numpy array
        self.global_step = 0
        self._experts = OrderedDict()      # variable prefix -> device.FcnExpert
        self._closed = False
        self._initialize_graph()

    # ------------------------------------------------------------------ graph lifecycle
    def _initialize_graph(self):
        """Counterpart of base_model.py:98-172: (re)build the experts."""
        dev.init()
        self._build_graph()
        if not self.custom_training and not hasattr(self, 'prediction'):
            self.prediction = 'prediction'
        self._cm_device = torch.zeros(
            (self.config['num_classes'], self.config['num_classes']), dtype=torch.int64,
            device='cuda')

    def _build_graph(self):
        raise NotImplementedError

    def _run_batch(self, batch, fetch='prediction'):
        """Runs one batch (dict of CUDA tensors) and returns the CUDA tensor named `fetch`."""
        raise NotImplementedError

    def _register_expert(self, prefix, expert, params):
        """Adds an expert and its variables (named `<prefix>/<layer>/<var>`)."""
        self._experts[prefix] = expert
        for name, value in params.items():
            self.variables[name] = np.asarray(value, dtype=np.float32)
        self._push_variables(prefix)

    def _push_variables(self, prefix=None):
        for pre, expert in self._experts.items():
            if prefix is not None and pre != prefix:
                continue
            head = pre + '/'
            expert.set_params({n[len(head):]: v for n, v in self.variables.items()
                               if n.startswith(head)})

    # ------------------------------------------------------------------ data plumbing
    def _my_rows(self, first_index, count):
        """Rows of a batch this rank evaluates (images sharded round-robin over ranks)."""
        rank, world = sharding.rank_world()
        if world == 1 or not self.config.get('shard_images', True):
            return slice(None)
        return sharding.rows_for_rank(first_index, count, rank, world)

    def _batches(self, data):
        """Replaces transform_inputdata (base_model.py:10-38): `data` is a dict of arrays
        batched on axis 0 (numpy or torch), or any iterable of such dicts.  Yields dicts of
        host/device arrays holding this rank's share of every batch."""
        if isinstance(data, dict):
            total = len(next(iter(data.values())))
            bs = self.config['batchsize']
            shard = self.config.get('shard_images', True)
            step = bs * (sharding.rank_world()[1] if shard else 1)

            def gen():
                for start in range(0, total, step):
                    stop = min(start + step, total)
                    rows = self._my_rows(start, stop - start)
                    if isinstance(rows, slice):
                        yield {k: v[start:stop] for k, v in data.items()}
                    elif rows:
                        idx = [start + r for r in rows]
                        yield {k: v[idx] for k, v in data.items()}
            return gen()

        def gen_iter():
            seen = 0
            for blob in data:
                count = len(next(iter(blob.values())))
                rows = self._my_rows(seen, count)
                seen += count
                if isinstance(rows, slice):
                    yield blob
                elif rows:
                    yield {k: v[rows] for k, v in blob.items()}
        return gen_iter()

    def _device_batches(self, data, presharded=False):
        """Yields this rank's batches as dicts of CUDA tensors.  Host batches are uploaded on
        a side stream one batch ahead, so the H2D copy of batch i+1 overlaps the kernels of
        batch i (the reference feeds every sess.run synchronously, base_model.py:308-313).
        presharded: `data` already yields this rank's batches (the training stream)."""
        compute = torch.cuda.current_stream()
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream()
        copy = self._copy_stream

        def upload(host_batch):
            if all(isinstance(v, torch.Tensor) and v.is_cuda for v in host_batch.values()):
                return self._to_device(host_batch), None
            copy.wait_stream(compute)          # never overwrite buffers the kernels still read
            tensors, events = {}, {}
            with torch.cuda.stream(copy):
                # smallest image array first (nothing can hide the first copy), labels last; one
                # event per array, so an expert starts as soon as ITS modality has arrived (see
                # _DeviceBatch and FusionModel._arrival_order)
                for key in upload_order(host_batch):
                    tensors.update(self._to_device({key: host_batch[key]}))
                    events[key] = torch.cuda.Event()
                    events[key].record(copy)
            for t in tensors.values():
                t.record_stream(compute)
            return _DeviceBatch(tensors, events, compute), None

        def pieces():
            # A large host batch is uploaded in pieces so that its own copy overlaps its own
            # kernels (matters when score() is called one batch at a time): [n/4, 3n/4] for
            # `upload_split` = 2 (the default), or the explicit sizes of `upload_pieces`.
            # Default schedule (measured, tools/e2e_split_probe.py): with two or more image
            # arrays the batch goes up whole, smallest array first - its expert runs while the
            # larger arrays are still in flight (3.20 k vs 3.03 k frames/s for [4, 12] pieces on
            # the RGB-D batch of 16); a single-modality batch is cut as [n/4, 3n/4] so that most
            # of its only copy overlaps its own kernels.
            split = self.config.get('upload_split')
            explicit = self.config.get('upload_pieces')      # e.g. [4, 12]: sizes of the pieces
            first = True
            for blob in (data if presharded else self._batches(data)):
                count = len(next(iter(blob.values())))
                on_host = not all(isinstance(v, torch.Tensor) and v.is_cuda
                                  for v in blob.values())
                # Only the FIRST batch of a data set is cut (nothing else can hide its copy);
                # the upload of every later batch already overlaps the kernels of the batch
                # before it, and a piece costs ~0.5 ms of small-batch inefficiency.  Training
                # batches are never cut: a piece would become an optimizer step.
                n_split = int(split) if split is not None else \
                    (1 if sum(1 for k in blob if k != 'labels') >= 2 else 2)
                bounds = (upload_bounds(count, n_split, explicit)
                          if on_host and first and not presharded else None)
                first = False
                if bounds is None:
                    yield blob
                    continue
                for start, stop in zip(bounds[:-1], bounds[1:]):
                    yield {k: v[start:stop] for k, v in blob.items()}

        it = iter(pieces())
        try:
            pending = upload(next(it))
        except StopIteration:
            return
        # The host may run at most `depth` batches ahead of the device: without a bound a long
        # data set is enqueued in one go, every upload allocates fresh device buffers (the
        # previous ones are still in flight) and the caching allocator falls back to cudaMalloc.
        in_flight = []
        depth = int(self.config.get('upload_depth', 3))
        while pending is not None:
            dev_batch, event = pending
            try:
                nxt = next(it)
            except StopIteration:
                nxt = None
            if event is not None:
                compute.wait_event(event)
            if len(in_flight) >= depth:
                in_flight.pop(0).synchronize()
            # issue the next upload before the caller launches this batch's kernels
            pending = upload(nxt) if nxt is not None else None
            yield dev_batch
            done = torch.cuda.Event()
            done.record(compute)              # the caller has enqueued this batch's kernels
            in_flight.append(done)

    @staticmethod
    def _to_device(batch):
        out = {}
        for key, value in batch.items():
            if isinstance(value, np.ndarray):
                value = torch.from_numpy(np.ascontiguousarray(value))
            # spatial size must be a multiple of 16 (crop_multiple, augmentation.py:244-262)
            value = crop_multiple(value, batched=True)
            if key == 'labels':
                if value.dim() == 4:        # one-hot labels of the training pipeline
                    # an all-zero row is what tf.one_hot builds for negative / void labels
                    # (base_model.py:198-201): it must stay "ignore", not become class 0
                    value = torch.where(value.sum(-1) > 0, value.argmax(-1),
                                        torch.full_like(value[..., 0], -1, dtype=torch.int64))
                out[key] = value.to(device='cuda', dtype=torch.int32, non_blocking=True)
            elif value.dtype in (torch.uint8, torch.uint16, torch.int16, torch.int32):
                # raw sensor dtype: copy the narrow values, cast to float32 on the device
                out[key] = dev.convert_to_f32(value.contiguous().to(device='cuda',
                                                                    non_blocking=True))
            else:
                out[key] = value.to(device='cuda', dtype=torch.float32, non_blocking=True)
        return out

    # ------------------------------------------------------------------ public API
    def fit(self, dataset, iterations, output=True, validation_dataset=None,
            validation_interval=100, additional_eval_datasets={}):
        """base_model.py:179-261.  Gradient training of the expert is outside the round-1
        scope of the B200 path."""
        raise UserWarning("ERROR: Model %s does not support training" % self.name)

    def predict(self, data, output_attr=None):
        """base_model.py:263-292: returns the concatenation over all batches of
        `self.prediction` ([num_images,H,W] int64) or of the attribute named `output_attr`."""
        fetch = 'prediction'
        if output_attr is not None and self._has_output(output_attr):
            fetch = output_attr
        ret = []
        for batch in self._device_batches(data):
            out = self._run_batch(batch, fetch)
            ret.append(out)
        local = torch.cat(ret) if ret else None
        # every rank evaluated its share of the images (its rows of the caller's dict, or with
        # shard_images=False the whole dict it was handed): interleave the shares back together.
        # Not so when the ranks split the MC SAMPLES of the same images (split_samples).
        if _dist() is not None and not self._same_images_on_every_rank():
            local = sharding.gather_interleaved(local, 'cuda')
        return local.cpu().numpy()

    # Small batches leave most of the 148 SMs idle in the deep layers (one 768x384 frame is nine
    # 128-pixel tiles in conv5_x), and the experts are independent until the fusion rule: up to
    # this many pixels per modality they run concurrently, each on its own CUDA stream
    # (config `overlap_experts`: True / False / None = by size).
    OVERLAP_MAX_PIXELS = 4 * 768 * 384

    def _overlap_experts(self, batch):
        choice = self.config.get('overlap_experts')
        if choice is not None:
            return bool(choice) and len(self.modalities) > 1
        if len(self.modalities) < 2 or not isinstance(batch, dict):
            return False
        x = dict.__getitem__(batch, self.modalities[0])
        return x.shape[0] * x.shape[1] * x.shape[2] <= self.OVERLAP_MAX_PIXELS

    def _run_experts(self, batch, fn, order=None):
        """{modality: fn(modality, input)} for every modality, experts in arrival order.  Below
        OVERLAP_MAX_PIXELS each call is enqueued on a side stream forked from the current one
        and joined before returning, so the caller keeps its single-stream view."""
        if order is None:      # smallest upload first (see upload_order)
            order = sorted(self.modalities, key=lambda m: _nbytes(dict.__getitem__(batch, m))
                           if isinstance(batch, dict) else 0)
        order = list(order)
        if not self._overlap_experts(batch):
            return {m: fn(m, batch[m]) for m in order}
        current = torch.cuda.current_stream()
        if not hasattr(self, '_side_streams'):
            self._side_streams = {}
        outputs, used = {}, []
        for m in order:
            x = batch[m]                   # the current stream waits for this modality's upload
            side = self._side_streams.get(m)
            if side is None:
                side = self._side_streams[m] = torch.cuda.Stream()
            side.wait_stream(current)
            with torch.cuda.stream(side):
                outputs[m] = fn(m, x)
            used.append(side)
        for side in used:
            current.wait_stream(side)
        return outputs

    def _same_images_on_every_rank(self):
        """True when the ranks cooperate on the SAME images (`split_samples`: each draws a share
        of the MC-dropout samples and the moments are merged inside the batch) - then every rank
        already ends with the complete result and nothing is gathered or reduced afterwards."""
        return bool(self.config.get('split_samples', False))

    def _has_output(self, attr):
        return attr in getattr(self, 'output_attrs', ('prediction',))

    def score(self, data, max_iterations=None):
        """base_model.py:294-331: returns (measures dict, confusion matrix float64 [C,C]).
        The confusion matrix is accumulated on the device in int64 and read back once."""
        cm = self._cm_device
        cm.zero_()
        for batch in self._device_batches(data):
            self._score_batch(batch, cm)
        if not self._same_images_on_every_rank():
            sharding.allreduce_sum_(cm)
        confusion_matrix = cm.cpu().numpy().astype(np.float64)
        measures = measures_from_confusion_matrix(confusion_matrix)
        return measures, confusion_matrix

    def score_batch_on_device(self, batch, confusion_matrix=None):
        """One step of the score() loop for a batch that already lives in HBM (dict of CUDA
        tensors): experts -> fusion -> confusion-matrix accumulation, no host transfer and no
        synchronisation.  Returns the int64 device matrix being accumulated."""
        cm = self._cm_device if confusion_matrix is None else confusion_matrix
        self._score_batch(batch, cm)
        return cm

    def capture_score_step(self, batch, confusion_matrix=None):
        """CUDA-graph form of score_batch_on_device for a batch that stays at the same device
        addresses (a resident evaluation set that is scored repeatedly): captures experts ->
        fusion -> confusion-matrix accumulation once and returns a callable that replays the
        whole step as ONE graph launch on the current stream.  The two warm-up passes and the
        capture itself do not change `confusion_matrix`'s meaning: it is zeroed afterwards."""
        cm = self._cm_device if confusion_matrix is None else confusion_matrix
        current = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(current)
        with torch.cuda.stream(side):
            for _ in range(2):                 # sizes the arenas / descriptor caches
                self._score_batch(batch, cm)
        current.wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        launches = dev.launch_count()
        with torch.cuda.graph(graph):
            self._score_batch(batch, cm)
        launches = dev.launch_count() - launches
        torch.cuda.synchronize()
        cm.zero_()
        self._graphs = getattr(self, '_graphs', []) + [graph]      # keep the pools alive
        return graph, launches

    def _score_batch(self, batch, cm):
        """One batch of score(): prediction + confusion-matrix accumulation into `cm` (device
        int64).  Models may override it with a fused path (BayesFusion does)."""
        prediction = self._run_batch(batch, 'prediction_compact')
        dev.confusion_accumulate(prediction, batch['labels'].contiguous(), cm)

    def load_weights(self, filepath):
        """base_model.py:333-339 restores a TensorFlow checkpoint; only the npz route exists
        here."""
        raise UserWarning('ERROR: TensorFlow checkpoints are not supported, use import_weights')

    def close(self):
        for expert in self._experts.values():
            expert.close()
        self._experts = OrderedDict()
        self._closed = True

    def __exit__(self, *args):
        self.close()

    def __enter__(self):
        return self

    def export_weights(self, save_dir=None):
        """base_model.py:361-393: every variable under its TensorFlow name, plus global_step,
        into `<name>_weights_<step>.npz`."""
        if save_dir is None and self.output_dir is None:
            print('ERROR: No path specified to save weights to.')
            return
        save_dict = {name: value for name, value in self.variables.items()}
        save_dict['global_step'] = np.asarray(self.global_step)
        output_path = save_dir if save_dir is not None else self.output_dir
        output_path = path.join(output_path, '{}_weights_{}.npz'.format(self.name,
                                                                        int(self.global_step)))
        np.savez_compressed(output_path, **save_dict)
        print('INFO: Weights saved to {}'.format(output_path))
        return output_path

    def import_weights(self, filepath, translate_prefix=False, chill_mode=False, warnings=True):
        """base_model.py:395-451: assign arrays of an npz file to the variables of equal name
        (or equal name with the first '/' replaced by '_', the legacy `rgb_conv1_1/kernel`
        style).  Optimizer slots are skipped, unknown names and shape mismatches only warn."""
        if warnings:
            print(filepath)
        weights = np.load(filepath)
        assigned, messages = match_weights(self.variables, weights, translate_prefix, chill_mode)
        if warnings:
            for line in messages:
                print(line)
        for var_name, value in assigned.items():
            self.variables[var_name] = value
        if 'global_step' in weights:
            self.global_step = int(weights['global_step'])
        self._push_variables()


def match_weights(variables, weights, translate_prefix=False, chill_mode=False):
    """The name-matching rules of base_model.py:409-450 as a pure function.

    variables: mapping TF variable name -> current array; weights: mapping stored name -> array
    (an opened npz).  Returns ({variable name: float32 array to assign}, [warning lines])."""
    keys = list(weights.keys())
    import_prefix = keys[0].split('/')[0].split('_')[0] if keys else ''

    def translate_name(name):
        """Name under which a variable is looked up in the file when the file was written by a
        model with another prefix: the text before the first '_' of the variable's scope is
        swapped for the file's prefix (base_model.py:414-428; 'forest...' scopes are exempt)."""
        if not translate_prefix or not name.startswith(translate_prefix):
            return name
        scope, sep, rest = name.partition('/')
        first, underscore, tail = scope.partition('_')
        if first == 'forest':
            return name
        return import_prefix + underscore + tail + sep + rest

    assigned, messages = {}, []
    for var_name in variables.keys():
        name = translate_name(var_name)
        # optimizers have their own variables, do not load these (base_model.py:433)
        if 'grad' in name or 'Adam' in name or 'RMS' in name:
            continue
        legacy = name.replace('/', '_', 1)
        if name in weights or legacy in weights:
            if legacy in weights:
                name = legacy
            value = weights[name]
            if tuple(variables[var_name].shape) != tuple(value.shape):
                messages.append('WARNING: wrong shape found for {}, but ignored in '
                                'chill mode'.format(name))
                messages.append('stored shape:  {} expected shape:  {}'.format(
                    tuple(value.shape), tuple(variables[var_name].shape)))
                # reference docstring: mismatching variables are left unassigned
                continue
            assigned[var_name] = np.asarray(value, dtype=np.float32)
        else:
            messages.append('WARNING: {} not found in saved weights'.format(name))
    return assigned, messages
