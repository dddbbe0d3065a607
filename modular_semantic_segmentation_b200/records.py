"""On-disk records either side of the hot path.

* `dump_expert_predictions` - the `predictions.npz` archive of experiments/ibcc_fusion.py:18-42:
  per-expert label maps on the measure and the test split plus both ground truths, the input of
  the external IBCC fusion.
* `ExperimentData` - reader of a stored experiment (sacred FileStorageObserver layout
  `run.json / info.json / config.json / cout.txt` in a folder or a `<id>.zip`), the non-database
  branch of experiments/utils.py:60-125, so that published runs (net_config, starting weights,
  confusion matrices of a fit) can be replayed through the models of this package.
* `decode_record` - experiments/utils.py:40-57 (`reverse_convert_datatypes`): jsonpickle-flattened
  numpy arrays / tuples back to python objects.
"""
import ast
import json
import os
import zipfile
from copy import deepcopy

import numpy as np


def decode_record(data):
    """Undo the json flattening of a sacred record (experiments/utils.py:40-57)."""
    if isinstance(data, dict):
        if len(data) == 1 and 'values' in data:
            return decode_record(data['values'])
        if len(data) == 1 and 'py/tuple' in data:
            return decode_record(data['py/tuple'])
        if data.get('py/object') == 'numpy.ndarray':
            if 'dtype' in data:
                return np.array(data['values'], dtype=data['dtype'])
            return np.array(data['values'])
        return {key: decode_record(value) for key, value in data.items()}
    if isinstance(data, list):
        return [decode_record(item) for item in data]
    if isinstance(data, str) and data[:1] == '[':
        # the reference eval()s such strings; only literals are accepted here
        try:
            return ast.literal_eval(data)
        except (ValueError, SyntaxError):
            return data
    return data


def encode_array(array):
    """numpy array -> the flattened form sacred stores in info.json."""
    array = np.asarray(array)
    return {'py/object': 'numpy.ndarray', 'dtype': str(array.dtype), 'values': array.tolist()}


STORAGE_ENV = 'XVIEW_EXPERIMENT_STORAGE_FOLDER'


def default_storage_folder():
    """Where stored experiments live when the caller names no folder: the environment variable
    XVIEW_EXPERIMENT_STORAGE_FOLDER (the role `settings.EXPERIMENT_STORAGE_FOLDER` plays at
    experiments/utils.py:76-78)."""
    folder = os.environ.get(STORAGE_ENV)
    if not folder:
        raise UserWarning('ERROR: no experiment storage folder: pass storage_folder or set %s'
                          % STORAGE_ENV)
    return folder


class ExperimentData(object):
    """Stored experiment `exp_id` below `storage_folder` (folder `<id>/` or archive `<id>.zip`);
    storage_folder defaults to default_storage_folder(), so `ExperimentData(exp_id)` works as at
    bayes_mix.py:143-147 / dirichlet_mix.py:60-66."""

    _PARTS = (('info', 'info.json'), ('config', 'config.json'))

    def __init__(self, exp_id, storage_folder=None):
        if storage_folder is None:
            storage_folder = default_storage_folder()
        exp_id = str(exp_id)
        entries = os.listdir(storage_folder)
        if exp_id in entries:
            self.exp_path = os.path.join(storage_folder, exp_id)
            self.artifacts = os.listdir(self.exp_path)
            read = lambda name: open(os.path.join(self.exp_path, name), 'rb').read()
        elif exp_id + '.zip' in entries:
            self.zipfile = os.path.join(storage_folder, exp_id + '.zip')
            archive = zipfile.ZipFile(self.zipfile)
            self.artifacts = archive.namelist()
            contents = {name: archive.read(name) for name in
                        ('run.json', 'info.json', 'config.json', 'cout.txt')
                        if name in self.artifacts}
            archive.close()
            read = contents.__getitem__
        else:
            raise UserWarning('Specified experiment %s not found.' % exp_id)
        record = json.loads(read('run.json').decode('utf8'))
        for key, name in self._PARTS:
            record[key] = json.loads(read(name).decode('utf8'))
        record['captured_out'] = read('cout.txt').decode('utf8', 'replace') \
            if 'cout.txt' in self.artifacts else ''
        self.record = record

    def get_record(self):
        return decode_record(deepcopy(self.record))

    def get_artifact(self, name):
        """File-like object of a stored output file."""
        if name not in self.artifacts:
            raise UserWarning('ERROR: Artifact {} not found'.format(name))
        if hasattr(self, 'exp_path'):
            return open(os.path.join(self.exp_path, name), 'rb')
        return zipfile.ZipFile(self.zipfile).open(name)

    def get_weights(self):
        """Path of the stored weights npz (experiments/utils.py:153-162)."""
        if not hasattr(self, 'exp_path'):
            raise UserWarning('cannot load weights out of zipfile, please extract first')
        filename = next(name for name in self.artifacts if 'weights' in name)
        return os.path.join(self.exp_path, filename)

    def get_confusion_matrices(self):
        """`info.confusion_matrices` of a fusion fit (experiments/bayes_fusion.py:39-44)."""
        stored = self.record['info']['confusion_matrices']
        return {key: np.array(decode_record(deepcopy(value))) for key, value in stored.items()}


def write_experiment(storage_folder, exp_id, config, info, run=None, captured_out='',
                     as_zip=False):
    """Store an experiment in the layout `ExperimentData` reads (FileStorageObserver's)."""
    def flatten(obj):
        if isinstance(obj, np.ndarray):
            return encode_array(obj)
        if isinstance(obj, dict):
            return {str(key): flatten(value) for key, value in obj.items()}
        if isinstance(obj, (list, tuple)):
            return [flatten(item) for item in obj]
        if isinstance(obj, np.generic):
            return obj.item()
        return obj

    files = {'run.json': json.dumps(flatten(run or {'status': 'COMPLETED'})),
             'info.json': json.dumps(flatten(info)),
             'config.json': json.dumps(flatten(config)),
             'cout.txt': captured_out}
    if as_zip:
        with zipfile.ZipFile(os.path.join(storage_folder, '%s.zip' % exp_id), 'w') as archive:
            for name, text in files.items():
                archive.writestr(name, text)
        return os.path.join(storage_folder, '%s.zip' % exp_id)
    folder = os.path.join(storage_folder, str(exp_id))
    os.makedirs(folder, exist_ok=True)
    for name, text in files.items():
        with open(os.path.join(folder, name), 'w') as f:
            f.write(text)
    return folder


def dump_expert_predictions(net_config, data_description, measure_set, test_set, save_to,
                            starting_weights=None):
    """Run every expert of `net_config['prefixes']` on both splits and store
    `predictions.npz` with keys measure_<expert>, test_<expert>, measure_gt, test_gt
    (experiments/ibcc_fusion.py:18-42).  Returns the archive path."""
    from .models import get_model
    model = get_model(net_config['expert_model'])
    predictions = {}
    for expert, prefix in net_config['prefixes'].items():
        model_config = deepcopy(net_config)
        model_config['modality'] = expert
        model_config['prefix'] = prefix
        if isinstance(model_config.get('num_channels'), dict):
            model_config['num_channels'] = model_config['num_channels'][expert]
        with model(data_description=data_description, **model_config) as net:
            if starting_weights is not None:
                net.import_weights(starting_weights[prefix])
            predictions['measure_%s' % expert] = net.predict(measure_set)
            predictions['test_%s' % expert] = net.predict(test_set)
    # predictions are cropped to multiples of 16 (crop_multiple, augmentation.py:244-262): the
    # ground truths must describe the same pixels
    from .input_pipeline import crop_multiple
    predictions['measure_gt'] = np.asarray(crop_multiple(np.asarray(measure_set['labels']),
                                                         batched=True))
    predictions['test_gt'] = np.asarray(crop_multiple(np.asarray(test_set['labels']),
                                                      batched=True))
    os.makedirs(save_to, exist_ok=True)
    outfile = os.path.join(save_to, 'predictions.npz')
    np.savez_compressed(outfile, **predictions)
    return outfile


def dump_bayes_insight(net, batches, save_to):
    """experiments/bayes_fusion.py:47-69 (`collect_data`): run `net.get_insight` over the
    batches and store predictions.npz, likelihoods.npz, conditionals.npz and probs.npz, one
    positional array (arr_0, arr_1, ...) per batch as np.savez_compressed(path, *list) does.
    Returns the four paths."""
    collected = {'probs': [], 'likelihoods': [], 'conditionals': [], 'predictions': []}
    for batch in batches:
        insight = net.get_insight(batch)
        collected['probs'].append(insight[0])
        collected['likelihoods'].append(insight[1])
        collected['conditionals'].append(insight[2])
        collected['predictions'].append(insight[3])
    os.makedirs(save_to, exist_ok=True)
    paths = []
    for name in ('predictions', 'likelihoods', 'conditionals', 'probs'):
        paths.append(os.path.join(save_to, name + '.npz'))
        np.savez_compressed(paths[-1], *collected[name])
    return paths
