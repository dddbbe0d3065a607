"""CPU tests of the host-side logic of the B200 package: C-ABI export list, variable layout,
npz name matching, Bayes/Dirichlet table construction, the float64 Dirichlet fit, measures."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, ROOT


@pytest.fixture(scope='module')
def built_lib():
    import __graft_entry__
    __graft_entry__.build()
    from modular_semantic_segmentation_b200 import _abi
    return _abi.load()


def test_library_exports_every_symbol_of_the_header(built_lib):
    declared = set()
    for name in ('xview_b200.h', 'xview_b200_measure.h'):
        header = open(os.path.join(ROOT, 'include', name)).read()
        found = set(re.findall(r'^(?:int|const char\*)\s+(xv_[a-z0-9_]+)\s*\(', header, re.M))
        assert found, name
        declared |= found
    assert len(declared) >= 28
    from modular_semantic_segmentation_b200 import _abi
    assert declared == set(_abi.PROTOTYPES), declared ^ set(_abi.PROTOTYPES)
    for name in declared:
        assert hasattr(built_lib, name), name
    assert built_lib.xv_abi_version() == 1


def test_product_fails_loudly_without_a_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from modular_semantic_segmentation_b200 import _abi, device
    with pytest.raises(_abi.XViewError):
        device.init()
    n = ctypes.c_int()
    assert built_lib.xv_device_sm_count(ctypes.byref(n)) != 0
    assert built_lib.xv_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'modular_semantic_segmentation_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r'^\s*(import|from)\s+oracle', src, re.M), f


def test_tools_do_not_import_the_oracle():
    """Outside tests/ only __graft_entry__.smoke() and bench.py's CPU legs may execute oracle/;
    the xview facade and the scripts under tools/ must not."""
    for folder in ('tools', 'xview'):
        for base, _, files in os.walk(os.path.join(ROOT, folder)):
            for f in files:
                if f.endswith('.py'):
                    src = open(os.path.join(base, f)).read()
                    assert not re.search(r'^\s*(import|from)\s+oracle', src, re.M), f


def test_variable_layout_matches_reference_key_list():
    from modular_semantic_segmentation_b200.models.simple_fcn import init_fcn_variables
    keys = json.load(open(os.path.join(GOLDEN, 'fcn_weight_keys.json')))
    for prefix, cin in (('rgb', 3), ('depth', 1)):
        v = init_fcn_variables(prefix, cin, 64, 12, rng=np.random.default_rng(0))
        assert list(v) == keys[prefix]
        shapes = oracle.fcn_param_shapes(prefix, cin, 64, 12)
        assert {k: a.shape for k, a in v.items()} == shapes
        np.testing.assert_array_equal(v[prefix + '/upscore/kernel'],
                                      oracle.bilinear_filter((16, 16, 64, 64)))
        assert not v[prefix + '/conv3_2/bias'].any()
    bn = init_fcn_variables('rgb', 3, 8, 5, batchnorm=True, rng=np.random.default_rng(0))
    assert 'rgb/conv1_1/moving_variance' in bn and 'rgb/upscore/gamma' in bn


def test_match_weights_follows_reference_rules():
    from modular_semantic_segmentation_b200.models.base_model import match_weights
    variables = {'rgb/conv1_1/kernel': np.zeros((3, 3, 3, 64), np.float32),
                 'rgb/conv1_1/bias': np.zeros(64, np.float32),
                 'rgb/conv1_2/kernel': np.zeros((3, 3, 64, 64), np.float32),
                 'rgb/score/kernel': np.zeros((1, 1, 64, 12), np.float32),
                 'rgb/conv1_1/kernel/Adam': np.zeros((3, 3, 3, 64), np.float32)}
    stored = {'rgb/conv1_1/kernel': np.ones((3, 3, 3, 64)),            # exact name
              'rgb_conv1_1/bias': np.full(64, 2.0),                     # legacy first '/' -> '_'
              'rgb/conv1_2/kernel': np.ones((3, 3, 32, 64)),            # wrong shape -> skipped
              'rgb/conv1_1/kernel/Adam': np.full((3, 3, 3, 64), 9.0),   # optimizer slot
              'global_step': np.asarray(7)}
    assigned, messages = match_weights(variables, stored)
    assert set(assigned) == {'rgb/conv1_1/kernel', 'rgb/conv1_1/bias'}
    assert assigned['rgb/conv1_1/bias'].dtype == np.float32 and assigned['rgb/conv1_1/bias'][0] == 2
    text = '\n'.join(messages)
    assert 'WARNING: wrong shape found for rgb/conv1_2/kernel' in text
    assert 'WARNING: rgb/score/kernel not found in saved weights' in text
    assert 'Adam' not in text
    # translate_prefix: variables of prefix 'depth' are filled from an 'rgb' file
    variables_d = {'depth/conv1_1/bias': np.zeros(64, np.float32)}
    assigned, _ = match_weights(variables_d, {'rgb/conv1_1/bias': np.full(64, 3.0)},
                                translate_prefix='depth')
    assert assigned['depth/conv1_1/bias'][0] == 3


def test_bayes_decision_table_reproduces_literal_rule(exp868):
    from modular_semantic_segmentation_b200.models.bayes_mix import (bayes_decision_matrix,
                                                                     bayes_decision_table,
                                                                     bayes_tables)
    cms = [exp868['cm_measure_rgb'].astype('float32').T,
           exp868['cm_measure_depth'].astype('float32').T]
    c = 12
    a, b = np.meshgrid(np.arange(c), np.arange(c), indexing='ij')
    for prior in ('data', 'uniform', 0.5):
        table = bayes_decision_table(cms, prior)
        ref = oracle.argmax_first(oracle.bayes_fusion([a, b], cms, prior)[0])
        np.testing.assert_array_equal(table, ref)
        log_cond, log_prior = bayes_tables(cms, prior)
        assert log_cond.dtype == np.float32 and log_cond.shape == (2, c, c)
    cms64 = [exp868['cm_measure_rgb'].T, exp868['cm_measure_depth'].T]
    np.testing.assert_array_equal(bayes_decision_matrix(cms64),
                                  oracle.bayes_decision_matrix(cms64))


def test_dirichlet_tables_match_oracle_terms():
    from modular_semantic_segmentation_b200.models.dirichlet_mix import (class_prior_from_counts,
                                                                         dirichlet_tables)
    rng = np.random.default_rng(0)
    c = 12
    params = [1 + rng.gamma(2, 2, size=(c, c)) for _ in range(2)]
    counts = rng.integers(0, 100, size=c)
    prior = class_prior_from_counts(counts, 'data')
    np.testing.assert_array_equal(prior, oracle.dirichlet_prior(counts))
    am1, lognorm, logprior = dirichlet_tables(params, 0.5, prior)
    assert am1.shape == (2, c, c) and lognorm.shape == (2, c) and logprior.shape == (c,)
    alpha = 0.5 * params[1].astype(np.float32).astype(np.float64)
    np.testing.assert_allclose(lognorm[1], oracle.dirichlet_log_norm(alpha), rtol=1e-6)
    # a pixel evaluated by hand equals the oracle score
    p = rng.dirichlet(np.ones(c), size=2).astype(np.float32)
    ref = oracle.dirichlet_fusion([p[0][None, None, None], p[1][None, None, None]], params, prior,
                                  sigma=0.5, dtype=np.float64)[0, 0, 0]
    mine = sum(np.log(1e-20 + (p[m] / p[m].sum()).astype(np.float64)) @ am1[m].astype(np.float64)
               - lognorm[m] for m in range(2)) + logprior
    np.testing.assert_allclose(mine, ref, rtol=1e-5, atol=1e-4)


def test_host_dirichlet_fit_matches_reference_outputs(dirichlet_golden):
    from modular_semantic_segmentation_b200.models.dirichletDifferentiation import \
        findDirichletPriors
    g = dirichlet_golden
    for i in g['fit_cases']:
        delta, beta = g['fit%d_delta_beta' % i]
        c = len(g['fit%d_ss' % i])
        alpha = findDirichletPriors(g['fit%d_ss' % i], g['fit%d_neg_ss' % i], np.ones(c),
                                    max_iter=10000, delta=delta, beta=beta)
        np.testing.assert_allclose(np.asarray(alpha, np.float64), g['fit%d_alpha' % i],
                                   rtol=1e-9)


def test_measures_match_stored_run(exp868):
    from modular_semantic_segmentation_b200.models.base_model import \
        measures_from_confusion_matrix
    m = measures_from_confusion_matrix(exp868['cm_test_fusion'])
    assert m['mean_IoU'] == 0.6877204624612501
    assert m['total_accuracy'] == 0.920302592013555
    assert set(m) == {'confusion_matrix', 'recall', 'precision', 'F1', 'mean_F1',
                      'total_accuracy', 'IoU', 'mean_IoU'}


def test_get_model_names():
    from xview.models import get_model
    import xview.models.bayes_mix as bm
    assert get_model('fcn').__name__ == 'SimpleFCN'
    assert get_model('bayes_fusion') is get_model('bayes_mix') is bm.BayesFusion
    assert get_model('dirichlet_mix').__name__ == 'DirichletFusion'
    assert get_model('average_fusion').__name__ == 'AverageFusion'
    assert get_model('variance_mix').__name__ == 'VarianceFusion'
    with pytest.raises(UserWarning):
        get_model('nope')


def test_input_pipeline_crop_and_collate():
    from modular_semantic_segmentation_b200.input_pipeline import collate, crop_multiple
    img = np.zeros((37, 50, 3), np.uint8)
    assert crop_multiple(img).shape == (32, 48, 3)
    assert crop_multiple(np.zeros((4, 37, 50)), batched=True).shape == (4, 32, 48)
    assert crop_multiple(np.zeros((32, 48, 1))).shape == (32, 48, 1)
    assert crop_multiple(5) == 5
    items = [{'rgb': np.full((33, 40, 3), i, np.uint8), 'depth': np.full((33, 40, 1), i, np.uint16),
              'labels': np.full((33, 40), i, np.int64)} for i in range(3)]
    batch = collate(items)
    assert batch['rgb'].shape == (3, 32, 32, 3) and batch['rgb'].dtype == np.uint8
    assert batch['depth'].dtype == np.uint16 and batch['labels'].dtype == np.int32
    assert collate(items, keep_raw_dtype=False)['rgb'].dtype == np.float32


def test_experiment_records_roundtrip(tmp_path):
    from modular_semantic_segmentation_b200.records import (ExperimentData, decode_record,
                                                           write_experiment)
    cms = {'rgb': np.arange(9.0).reshape(3, 3), 'depth': np.eye(3)}
    config = {'net_config': {'num_units': 8, 'prefixes': {'rgb': 'rgb', 'depth': 'depth'}},
              'starting_weights': {'rgb': 12, 'depth': 13}}
    for as_zip, exp_id in ((False, 868), (True, 869)):
        write_experiment(str(tmp_path), exp_id, config, {'confusion_matrices': cms},
                         captured_out='done', as_zip=as_zip)
        exp = ExperimentData(exp_id, str(tmp_path))
        record = exp.get_record()
        assert record['config']['net_config']['num_units'] == 8
        assert record['captured_out'] == 'done'
        got = exp.get_confusion_matrices()
        np.testing.assert_array_equal(got['rgb'], cms['rgb'])
        np.testing.assert_array_equal(record['info']['confusion_matrices']['depth'], cms['depth'])
        assert exp.get_artifact('cout.txt').read() == b'done'
    with pytest.raises(UserWarning):
        ExperimentData(1, str(tmp_path))
    with pytest.raises(UserWarning):
        ExperimentData(869, str(tmp_path)).get_weights()
    assert decode_record({'py/tuple': [1, 2]}) == [1, 2]
    assert decode_record('[1, 2]') == [1, 2]
    assert decode_record({'a': {'values': [3]}}) == {'a': [3]}


def test_adapnet_variable_layout_matches_oracle():
    from modular_semantic_segmentation_b200.models.adapnet import init_adapnet_variables
    mine = init_adapnet_variables('depth', 1, 20, 14, np.random.default_rng(0))
    shapes = oracle.adapnet_param_shapes('depth', 1, 20, 14)
    assert {k: v.shape for k, v in mine.items()} == {k: tuple(s) for k, s in shapes.items()}
    up = mine['depth/second_deconvolution_upconv/kernel']
    assert up.shape == (16, 16, 14, 20) and up[:, :, 3, 3].max() > 0 and up[:, :, 3, 4].max() == 0


def test_upload_order_smallest_image_first_labels_last():
    from modular_semantic_segmentation_b200.models.base_model import upload_order
    batch = {'rgb': np.zeros((2, 16, 16, 3), np.float32), 'labels': np.zeros((2, 16, 16), np.int32),
             'depth': np.zeros((2, 16, 16, 1), np.float32)}
    assert upload_order(batch) == ['depth', 'rgb', 'labels']
    assert upload_order({'labels': batch['labels'], 'rgb': batch['rgb']}) == ['rgb', 'labels']


def test_upload_bounds():
    from modular_semantic_segmentation_b200.models.base_model import upload_bounds
    assert upload_bounds(16) == [0, 4, 16]            # a quarter first, then the rest
    assert upload_bounds(32, 2) == [0, 8, 32]
    assert upload_bounds(24, 3) == [0, 4, 14, 24]
    assert upload_bounds(15) is None                  # small batches go up in one piece
    assert upload_bounds(16, 1) is None
    assert upload_bounds(16, 2, [2, 6, 8]) == [0, 2, 8, 16]
    assert upload_bounds(16, 2, [4, 4]) == [0, 4, 16]     # sizes that do not add up are ignored
    for n in range(16, 70):
        b = upload_bounds(n)
        assert b[0] == 0 and b[-1] == n and all(x < y for x, y in zip(b, b[1:]))


def test_device_batch_waits_per_array():
    """_DeviceBatch: reading one entry makes the compute stream wait for that entry's upload
    only; iterating waits for everything; nothing is waited for twice."""
    from modular_semantic_segmentation_b200.models.base_model import _DeviceBatch

    class Stream(object):
        def __init__(self):
            self.waited = []

        def wait_event(self, event):
            self.waited.append(event)

    stream = Stream()
    batch = _DeviceBatch({'rgb': 1, 'depth': 2, 'labels': 3},
                         {'rgb': 'e_rgb', 'depth': 'e_depth', 'labels': 'e_labels'}, stream)
    assert 'rgb' in batch and len(batch) == 3 and stream.waited == []
    assert batch['rgb'] == 1 and stream.waited == ['e_rgb']
    assert batch['rgb'] == 1 and stream.waited == ['e_rgb']
    assert batch.get('depth') == 2 and stream.waited == ['e_rgb', 'e_depth']
    assert batch.get('missing', 7) == 7
    assert sorted(batch.values()) == [1, 2, 3]
    assert stream.waited == ['e_rgb', 'e_depth', 'e_labels']
    assert dict(batch.items()) == {'rgb': 1, 'depth': 2, 'labels': 3}
    assert stream.waited == ['e_rgb', 'e_depth', 'e_labels']


def test_dump_expert_predictions_layout(tmp_path, monkeypatch):
    """predictions.npz of experiments/ibcc_fusion.py:18-42 with a stand-in expert model (the GPU
    version of this test runs the real experts)."""
    from modular_semantic_segmentation_b200 import models, records

    class FakeExpert(object):
        def __init__(self, data_description=None, **config):
            self.offset = {'rgb': 1, 'depth': 2}[config['modality']]
            assert config['prefix'] == 'p_' + config['modality']

        def __enter__(self):
            return self

        def __exit__(self, *args):
            return False

        def import_weights(self, path):
            self.path = path

        def predict(self, data):
            # the real models crop every input to multiples of 16 (crop_multiple)
            return (np.asarray(data['labels'])[:, :16, :16] + self.offset).astype(np.int64)

    monkeypatch.setattr(models, 'get_model', lambda name: FakeExpert)
    measure = {'labels': np.zeros((3, 16, 16), np.int32)}
    test = {'labels': np.ones((2, 20, 18), np.int32)}
    out = records.dump_expert_predictions(
        {'expert_model': 'fcn', 'prefixes': {'rgb': 'p_rgb', 'depth': 'p_depth'}}, None, measure,
        test, str(tmp_path / 'out'), starting_weights={'p_rgb': 'a.npz', 'p_depth': 'b.npz'})
    with np.load(out) as archive:
        assert sorted(archive.files) == ['measure_depth', 'measure_gt', 'measure_rgb',
                                         'test_depth', 'test_gt', 'test_rgb']
        assert (archive['measure_rgb'] == 1).all() and (archive['test_depth'] == 3).all()
        # ground truths are cropped like the predictions (same pixels on both sides)
        np.testing.assert_array_equal(archive['test_gt'], test['labels'][:, :16, :16])
        assert archive['test_gt'].shape == archive['test_rgb'].shape
        np.testing.assert_array_equal(archive['measure_gt'], measure['labels'])


def test_bayes_insight_dump_layout(tmp_path):
    """records.dump_bayes_insight: four archives with one positional array per batch, the layout
    np.savez_compressed(path, *list) gives experiments/bayes_fusion.py:62-69."""
    from modular_semantic_segmentation_b200 import records

    class Stub(object):
        def get_insight(self, batch):
            n = len(batch['rgb'])
            return (np.full((2, n, 4, 4, 3), 1 / 3.0), np.zeros((2, n, 4, 4, 3)),
                    np.ones((2, n, 4, 4, 3)), np.arange(n * 16).reshape(n, 4, 4))

    batches = [{'rgb': np.zeros((2, 4, 4, 3))}, {'rgb': np.zeros((1, 4, 4, 3))}]
    paths = records.dump_bayes_insight(Stub(), batches, str(tmp_path / 'out'))
    assert [os.path.basename(p) for p in paths] == ['predictions.npz', 'likelihoods.npz',
                                                    'conditionals.npz', 'probs.npz']
    predictions = np.load(paths[0])
    assert predictions.files == ['arr_0', 'arr_1'] and predictions['arr_1'].shape == (1, 4, 4)
    assert np.load(paths[3])['arr_0'].shape == (2, 2, 4, 4, 3)


def test_clock_sampler_prefers_nvml_rows_and_falls_back_to_nvidia_smi_rows():
    """bench.ClockSampler.stop: samples inside the timed region, NVML first, the nvidia-smi
    witness when NVML delivered nothing (seen on one box), widened by 50 ms when fewer than five
    fall inside."""
    import bench
    sampler = bench.ClockSampler.__new__(bench.ClockSampler)
    sampler.proc, sampler._stop, sampler.source, sampler._index = None, False, 'nvml', 0
    t0 = 1000.0
    sampler.rows = [(t0 + 0.01 * i, 1500.0 + i, 1965.0, ['sw_power_cap']) for i in range(10)]
    sampler.smi_rows = [(t0 + 0.02, 1400.0, 1965.0, [])]
    got = sampler.stop(t0, t0 + 0.1)
    assert got['source'] == 'nvml' and got['samples'] == 10 and got['reasons'] == ['sw_power_cap']
    assert got['sm_mhz'] == 1504.5 and got['sm_min_mhz'] == 1500.0 and got['sm_max_mhz'] == 1965.0
    sampler.rows = []
    got = sampler.stop(t0, t0 + 0.1)
    assert got['source'] == 'nvidia-smi' and got['samples'] == 1 and got['sm_mhz'] == 1400.0
    sampler.smi_rows = [(t0 - 0.03, 1300.0, 1965.0, ['hw_slowdown'])]       # only in the margin
    got = sampler.stop(t0, t0 + 0.1)
    assert got['samples'] == 1 and got['reasons'] == ['hw_slowdown']


def test_small_batches_overlap_their_experts():
    """BaseModel._overlap_experts: by size (at most four 768x384 frames) unless the config
    decides; never for a single modality."""
    from modular_semantic_segmentation_b200.models.base_model import BaseModel
    model = BaseModel.__new__(BaseModel)
    model.modalities = ['rgb', 'depth']
    model.config = {}
    small = {'rgb': np.zeros((4, 768, 384, 3), np.float32), 'depth': np.zeros((4, 768, 384, 1), np.float32)}
    large = {'rgb': np.zeros((5, 768, 384, 3), np.uint8), 'depth': np.zeros((5, 768, 384, 1), np.uint8)}
    assert model._overlap_experts(small) and not model._overlap_experts(large)
    model.config = {'overlap_experts': True}
    assert model._overlap_experts(large)
    model.config = {'overlap_experts': False}
    assert not model._overlap_experts(small)
    model.config, model.modalities = {'overlap_experts': True}, ['rgb']
    assert not model._overlap_experts(small)
    model.config = {'split_samples': True}
    assert model._same_images_on_every_rank()
