"""GPU parity: the model classes (the reference-facing API) against the oracle."""
import os

import numpy as np
import pytest
import torch

import oracle
from util import assert_labels_match, top2_margin

pytestmark = pytest.mark.gpu

NU = 8


def _description(c):
    return ({'rgb': np.float32, 'depth': np.float32, 'labels': np.int32},
            {'rgb': (None, None, 3), 'depth': (None, None, 1), 'labels': (None, None)}, c)


def _data(rng, n, h, w, c):
    return {'rgb': rng.integers(0, 256, size=(n, h, w, 3)).astype(np.float32),
            'depth': rng.integers(0, 65536, size=(n, h, w, 1)).astype(np.float32),
            'labels': rng.integers(-1, c, size=(n, h, w)).astype(np.int32)}


def _trained_like(rng, c, nu=NU):
    params = {}
    for m, cin, hi in (('rgb', 3, 255.0), ('depth', 1, 65535.0)):
        p = oracle.glorot_fcn_params(m, cin, nu, c, rng, gain=1.45, bias_scale=0.05)
        p[m + '/conv1_1/kernel'] /= np.float32(hi)
        params.update(p)
    return params


def _load(net, params):
    for name, value in params.items():
        if name in net.variables:
            net.variables[name] = value
    net._push_variables()


def test_simple_fcn_predict_and_score_config0(tmp_path):
    """BASELINE configs[0]: SimpleFCN rgb-only (README example: num_classes=10,
    dropout_probability=0.2 - a key the model ignores) predict + score on a Synthia-shaped
    batch, against the CPU oracle."""
    from xview.models import get_model
    c, n, h, w = 10, 2, 368, 640
    rng = np.random.default_rng(0)
    data = _data(rng, n, h, w, c)
    params = {k: v for k, v in _trained_like(rng, c, 16).items() if k.startswith('rgb/')}
    with get_model('fcn')('rgb', _description(c), 'rgb', num_units=16, batch_normalization=False,
                          dropout_probability=0.2, batchsize=1, output_dir=str(tmp_path)) as net:
        _load(net, params)
        pred = net.predict({'rgb': data['rgb']})          # labels may be omitted (README:79)
        measures, cm = net.score(data)
        prob = net.predict(data, output_attr='prob')
        path = net.export_weights()
    ref = oracle.test_pipeline(data['rgb'], params, 'rgb', 16, c)
    assert pred.dtype == np.int64 and pred.shape == (n, h, w)
    np.testing.assert_allclose(prob, ref['prob'], rtol=0, atol=2e-2)
    assert (pred == ref['classification']).mean() > 0.97
    # score() on the device's own predictions is exact integer work
    np.testing.assert_array_equal(cm, oracle.confusion_matrix(data['labels'], pred, c))
    ref_m = oracle.score_measures(cm)
    for key in ('mean_IoU', 'total_accuracy', 'mean_F1'):
        assert measures[key] == ref_m[key]
    # mIoU within 0.1 point of the oracle's end-to-end value
    ref_cm = oracle.confusion_matrix(data['labels'], ref['classification'], c)
    assert abs(measures['mean_IoU'] - oracle.score_measures(ref_cm)['mean_IoU']) < 1e-3
    assert os.path.basename(path) == 'SimpleFCN_weights_0.npz'
    stored = np.load(path)
    assert set(stored.keys()) == set(params) | {'global_step'}


def test_import_weights_roundtrip_and_legacy_names(tmp_path, capsys):
    from xview.models import get_model
    c = 5
    rng = np.random.default_rng(1)
    params = {k: v for k, v in _trained_like(rng, c).items() if k.startswith('depth/')}
    legacy = {k.replace('/', '_', 1): v for k, v in params.items()}     # depth_conv1_1/kernel
    legacy['depth_conv1_1/kernel/Adam'] = np.zeros((3, 3, 1, 64), np.float32)
    legacy.pop('depth_score/bias')
    np.savez_compressed(tmp_path / 'w.npz', **legacy)
    x = rng.integers(0, 65536, size=(1, 32, 32, 1)).astype(np.float32)
    with get_model('fcn')('depth', _description(c), 'depth', num_units=NU,
                          batch_normalization=False) as net:
        before = net.predict({'depth': x}, output_attr='score')
        net.import_weights(str(tmp_path / 'w.npz'))
        after = net.predict({'depth': x}, output_attr='score')
        np.testing.assert_array_equal(net.variables['depth/conv3_2/kernel'],
                                      params['depth/conv3_2/kernel'])
        assert not net.variables['depth/score/bias'].any()     # missing -> left at its init
    out = capsys.readouterr().out
    assert 'WARNING: depth/score/bias not found in saved weights' in out
    assert not np.allclose(before, after)
    p2 = dict(params)
    p2['depth/score/bias'] = np.zeros(c, np.float32)
    ref = oracle.fcn(x, p2, 'depth', NU, c)['score']
    np.testing.assert_allclose(after, ref, rtol=0, atol=0.05 * np.abs(ref).max())


@pytest.mark.parametrize('prior', ['data', 'uniform'])
def test_bayes_fusion_model(prior):
    from xview.models import get_model
    c, n, h, w = 6, 5, 32, 48
    rng = np.random.default_rng(2)
    data = _data(rng, n, h, w, c)
    params = _trained_like(rng, c)
    cms = {m: rng.integers(0, 60, size=(c, c)).astype(np.float64) + 150 * np.eye(c)
           for m in ('rgb', 'depth')}
    with get_model('bayes_fusion')(
            confusion_matrices=cms, data_description=_description(c),
            prefixes={'rgb': 'rgb', 'depth': 'depth'}, expert_model='fcn', num_units=NU,
            num_channels={'rgb': 3, 'depth': 1}, batchsize=2, class_prior=prior) as net:
        _load(net, params)
        fused = net.predict({'rgb': data['rgb'], 'depth': data['depth']})
        measures, cm = net.score(data)                      # 5 images in batches of 2: ragged
        score = net.predict(data, output_attr='fused_score')
        # one-image expert classifications of the last batch for the integer-parity check
        last = {m: net.expert_outputs[m]['classification'].cpu().numpy() for m in net.modalities}
    assert fused.shape == (n, h, w) and fused.dtype == np.int64
    tables = [cms[m].astype('float32').T for m in ('rgb', 'depth')]
    # fusion of the device's own expert labels: bit-exact vs the oracle's float32 rule
    ref_score, _, _ = oracle.bayes_fusion([last['rgb'].astype(np.int64),
                                           last['depth'].astype(np.int64)], tables, prior)
    np.testing.assert_array_equal(score[-1:], ref_score.astype(np.float32))
    np.testing.assert_array_equal(fused[-1:], oracle.argmax_first(ref_score))
    np.testing.assert_array_equal(cm, oracle.confusion_matrix(data['labels'], fused, c))
    # end to end vs the fp32 oracle experts: high agreement, mIoU close
    cls = [oracle.test_pipeline(data[m], params, m, NU, c)['classification']
           for m in ('rgb', 'depth')]
    ref_fused = oracle.argmax_first(oracle.bayes_fusion(cls, tables, prior)[0])
    assert (fused == ref_fused).mean() > 0.95
    ref_m = oracle.score_measures(oracle.confusion_matrix(data['labels'], ref_fused, c))
    assert abs(measures['mean_IoU'] - ref_m['mean_IoU']) < 1e-2


def test_dirichlet_fusion_fit_and_predict():
    from xview.models import get_model
    c, n, h, w = 5, 4, 32, 48
    rng = np.random.default_rng(3)
    data = _data(rng, n, h, w, c)
    data['labels'][data['labels'] == 3] = 1          # class 3 never occurs -> alpha column of ones
    params = _trained_like(rng, c)
    config = dict(data_description=_description(c), modalities=['rgb', 'depth'],
                  expert_model='fcn', num_units=NU, num_channels={'rgb': 3, 'depth': 1},
                  batchsize=2, sigma=0.8, delta=1e-2, beta=1e-2, class_prior='data')
    with get_model('dirichlet_fusion')(**config) as net:
        _load(net, params)
        with pytest.raises(UserWarning):
            net.predict(data)
        fit = net.fit(data)
        probs = {m: net._experts[m].forward(torch.from_numpy(data[m]).cuda(),
                                            want=('prob',))['prob'].cpu().numpy()
                 for m in ('rgb', 'depth')}
        fused = net.predict(data)
        score = net.predict(data, output_attr='fused_score')
    # sufficient statistics + host fit reproduce the oracle given the device's probabilities
    for m in ('rgb', 'depth'):
        s_ref, n_ref = oracle.sufficient_statistics(probs[m], data['labels'], c)
        alpha_ref = oracle.fit_sufficient_statistic(s_ref, n_ref, delta=1e-2, beta=1e-2)
        np.testing.assert_array_equal(fit['class_counts'], n_ref)
        np.testing.assert_allclose(fit[m], alpha_ref, rtol=2e-4)
        np.testing.assert_array_equal(fit[m][:, 3], np.ones(c))
    prior = oracle.dirichlet_prior(fit['class_counts'])
    ref = oracle.dirichlet_fusion([probs['rgb'], probs['depth']], [fit['rgb'], fit['depth']],
                                  prior, sigma=0.8, dtype=np.float64)
    scale = np.abs(ref).max()
    np.testing.assert_allclose(score, ref, rtol=0, atol=3e-5 * scale)
    assert_labels_match(fused, ref, 6e-5 * scale)
    # the model runs the exact fusion mode by default: scores and labels are bit-exact against
    # the fixed-order float32 statement of the rule
    ref32 = oracle.dirichlet_fusion_f32([probs['rgb'], probs['depth']],
                                        [fit['rgb'], fit['depth']], prior, sigma=0.8)
    np.testing.assert_array_equal(score, ref32)
    np.testing.assert_array_equal(fused, oracle.argmax_first(ref32))
    # a second model constructed from the fitted parameters gives the same prediction
    with get_model('dirichlet_mix')(dirichlet_params=fit, **config) as net2:
        _load(net2, params)
        np.testing.assert_array_equal(net2.predict(data), fused)


def test_average_and_variance_fusion_models():
    from xview.models import get_model
    c, n, h, w = 5, 2, 32, 32
    rng = np.random.default_rng(4)
    data = _data(rng, n, h, w, c)
    params = _trained_like(rng, c)
    common = dict(data_description=_description(c), prefixes={'rgb': 'rgb', 'depth': 'depth'},
                  expert_model='fcn', num_units=NU, num_channels={'rgb': 3, 'depth': 1},
                  batchsize=2)
    with get_model('average_fusion')(**common) as net:
        _load(net, params)
        fused = net.predict(data)
        probs = [net.expert_outputs[m]['prob'].cpu().numpy() for m in net.modalities]
    assert_labels_match(fused, oracle.average_fusion(probs), 1e-6)
    with get_model('variance_fusion')(modalities=['rgb', 'depth'], dropout_rate=0.3,
                                      num_samples=6, deterministic_dropout=True, seed=5,
                                      **common) as net:
        _load(net, params)
        fused = net.predict(data)
        score = net.predict(data, output_attr='fused_score')
        # the pieces the model fuses, recomputed through the expert API with the model's seeds:
        # dropout-free probabilities and the MC-dropout variance (variance_mix.py:62-69)
        from modular_semantic_segmentation_b200.models.variance_mix import mc_dropout_seed
        probs, variances = [], []
        for i, m in enumerate(net.modalities):
            x = torch.from_numpy(data[m]).cuda()
            expert = net._experts[m]
            probs.append(expert.forward(x, want=('prob',))['prob'].cpu().numpy())
            mc = expert.forward(x, want=('mean_var',), dropout={
                'rate': 0.3, 'layers': ['pool3'], 'num_samples': 6,
                'seed': mc_dropout_seed(net, i)})
            variances.append(mc['mean_var'].cpu().numpy()[..., None])
    assert fused.shape == (n, h, w)
    np.testing.assert_allclose(score.sum(-1), 1.0, atol=1e-4)   # convex mix of probabilities
    np.testing.assert_array_equal(fused, np.argmax(score, -1))
    # the shared-trunk call equals the two separate passes, and the fusion rule is bit-exact
    ref = oracle.variance_fusion(probs, variances)
    np.testing.assert_array_equal(score, ref)
    np.testing.assert_array_equal(fused, oracle.argmax_first(ref))
    # without `deterministic_dropout` every call draws fresh masks (the reference's behaviour)
    with get_model('variance_fusion')(modalities=['rgb', 'depth'], dropout_rate=0.3,
                                      num_samples=6, seed=5, **common) as net:
        _load(net, params)
        a = net.predict(data, output_attr='fused_score')
        b = net.predict(data, output_attr='fused_score')
    assert not np.array_equal(a, b)


def test_fusion_fcn_mid_level_fusion_net():
    """fusion_fcn (two VGG16 towers, concat of conv4_3 / conv5_3) against
    the oracle, through the model class, the functional API and the legacy variable names."""
    from xview.models import get_model
    from xview.models.fusion_fcn import fusion_fcn
    c, n, h, w = 6, 2, 48, 64
    rng = np.random.default_rng(8)
    data = _data(rng, n, h, w, c)
    prefixes = {'rgb': 'rgb', 'depth': 'depth'}
    channels = {'rgb': 3, 'depth': 1}
    params = oracle.fusion_fcn_params(prefixes, channels, NU, c, rng, gain=1.45, bias_scale=0.05)
    params['rgb_conv1_1/kernel'] /= np.float32(255.0)
    params['depth_conv1_1/kernel'] /= np.float32(65535.0)
    ref = oracle.fusion_fcn(data, params, prefixes, NU, c)
    ref_prob = oracle.softmax(ref['score'])
    with get_model('fusion_fcn')(_description(c), prefixes, channels, NU, batchsize=2) as net:
        assert set(net.variables) == set(params)
        _load(net, params)
        prob = net.predict(data, output_attr='prob')
        pred = net.predict({'rgb': data['rgb'], 'depth': data['depth']})
        measures, cm = net.score(data)
    np.testing.assert_allclose(prob, ref_prob, rtol=0, atol=2e-2)
    assert (pred == np.argmax(ref_prob, -1)).mean() > 0.97
    np.testing.assert_array_equal(cm, oracle.confusion_matrix(data['labels'], pred, c))
    out = fusion_fcn({m: torch.from_numpy(data[m]).cuda() for m in prefixes}, prefixes, NU, c,
                     want=('score',), params=params)
    np.testing.assert_allclose(out['score'].cpu().numpy(), ref['score'], rtol=0,
                               atol=0.05 * np.abs(ref['score']).max())


def test_raw_sensor_dtypes_and_crop_multiple():
    """uint8 rgb / uint16 depth inputs of arbitrary size are cropped to
    multiples of 16 and cast on the device; results equal the float32 path exactly."""
    from xview.models import get_model
    c, n, h, w = 6, 3, 37, 52
    rng = np.random.default_rng(12)
    raw = {'rgb': rng.integers(0, 256, size=(n, h, w, 3)).astype(np.uint8),
           'depth': rng.integers(0, 65536, size=(n, h, w, 1)).astype(np.uint16),
           'labels': rng.integers(-1, c, size=(n, h, w)).astype(np.int32)}
    as_float = {'rgb': raw['rgb'][:, :32, :48].astype(np.float32),
                'depth': raw['depth'][:, :32, :48].astype(np.float32),
                'labels': raw['labels'][:, :32, :48]}
    cms = {m: rng.integers(1, 50, size=(c, c)).astype(np.float64) + 100 * np.eye(c)
           for m in ('rgb', 'depth')}
    with get_model('bayes_fusion')(
            confusion_matrices=cms, data_description=_description(c),
            prefixes={'rgb': 'rgb', 'depth': 'depth'}, expert_model='fcn', num_units=NU,
            num_channels={'rgb': 3, 'depth': 1}, batchsize=2, seed=4) as net:
        a = net.predict({'rgb': raw['rgb'], 'depth': raw['depth']})
        b = net.predict({'rgb': as_float['rgb'], 'depth': as_float['depth']})
        _, cm_a = net.score(raw)
        _, cm_b = net.score(as_float)
    assert a.shape == (n, 32, 48)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(cm_a, cm_b)


def test_expert_predictions_dump(tmp_path):
    """predictions.npz of experiments/ibcc_fusion.py:18-42."""
    from modular_semantic_segmentation_b200.records import dump_expert_predictions
    c, h, w = 5, 32, 32
    rng = np.random.default_rng(2)

    def split(n):
        return {'rgb': rng.normal(size=(n, h, w, 3)).astype(np.float32),
                'depth': rng.normal(size=(n, h, w, 1)).astype(np.float32),
                'labels': rng.integers(0, c, size=(n, h, w)).astype(np.int32)}
    measure, test = split(3), split(2)
    net_config = {'expert_model': 'fcn', 'prefixes': {'rgb': 'rgb', 'depth': 'depth'},
                  'num_units': NU, 'num_channels': {'rgb': 3, 'depth': 1}, 'batchsize': 2,
                  'seed': 5, 'batch_normalization': False}
    out = dump_expert_predictions(net_config, _description(c), measure, test, str(tmp_path))
    with np.load(out) as archive:
        assert sorted(archive.files) == ['measure_depth', 'measure_gt', 'measure_rgb',
                                         'test_depth', 'test_gt', 'test_rgb']
        assert archive['measure_rgb'].shape == (3, h, w) and archive['test_depth'].shape == (2, h, w)
        assert archive['measure_rgb'].dtype == np.int64
        np.testing.assert_array_equal(archive['test_gt'], test['labels'])


def test_fusion_models_read_stored_experiments(tmp_path, monkeypatch):
    """bayes_mix.py:143-147 (`eval_experiments`) and dirichlet_mix.py:60-66 (`measurement_exp`):
    confusion matrices / Dirichlet measurements come from stored runs in the file layout of
    experiments/utils.py:79-104."""
    from xview.models import get_model
    from modular_semantic_segmentation_b200 import records
    c, n, h, w = 5, 2, 32, 32
    rng = np.random.default_rng(31)
    data = _data(rng, n, h, w, c)
    params = _trained_like(rng, c)
    cms = {m: rng.integers(0, 60, size=(c, c)).astype(np.float64) + 150 * np.eye(c)
           for m in ('rgb', 'depth')}
    for i, m in enumerate(('rgb', 'depth')):
        records.write_experiment(str(tmp_path), 100 + i, {'modality': m},
                                 {'confusion_matrix': cms[m]}, as_zip=bool(i))
    common = dict(data_description=_description(c), prefixes={'rgb': 'rgb', 'depth': 'depth'},
                  expert_model='fcn', num_units=NU, num_channels={'rgb': 3, 'depth': 1},
                  batchsize=2)
    with get_model('bayes_fusion')(confusion_matrices=cms, **common) as net:
        _load(net, params)
        want = net.predict(data)
    monkeypatch.setenv(records.STORAGE_ENV, str(tmp_path))
    with get_model('bayes_fusion')(eval_experiments={'rgb': 100, 'depth': 101}, **common) as net:
        assert net.modalities == ['rgb', 'depth']
        _load(net, params)
        np.testing.assert_array_equal(net.predict(data), want)
    # Dirichlet measurements stored as the counts.npz artifact of a measurement run
    fit = {m: 1.0 + rng.gamma(2.0, 2.0, size=(c, c)) for m in ('rgb', 'depth')}
    fit['class_counts'] = rng.integers(10, 1000, size=c).astype(np.float64)
    folder = records.write_experiment(str(tmp_path), 200, {}, {})
    np.savez(os.path.join(folder, 'counts.npz'), **fit)
    dcommon = dict(common, modalities=['rgb', 'depth'])
    dcommon.pop('prefixes')
    with get_model('dirichlet_mix')(dirichlet_params=fit, **dcommon) as net:
        _load(net, params)
        want = net.predict(data)
    with get_model('dirichlet_mix')(measurement_exp=200, experiment_storage_folder=str(tmp_path),
                                    **dcommon) as net:
        _load(net, params)
        np.testing.assert_array_equal(net.predict(data), want)


def test_custom_layers_softmax_log_softmax_entropy():
    """custom_layers.py:222-256 on the device against their numpy statements."""
    from modular_semantic_segmentation_b200.models import custom_layers as cl
    rng = np.random.default_rng(17)
    x = rng.normal(0, 3, size=(2, 9, 11, 7)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    ref = oracle.softmax(x)
    np.testing.assert_allclose(cl.softmax(xd).cpu().numpy(), ref, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(cl.softmax(xd, temperature=2.0).cpu().numpy(),
                               oracle.softmax(x / 2), rtol=2e-6, atol=1e-7)
    d = x - x.max(-1, keepdims=True)
    np.testing.assert_allclose(cl.log_softmax(xd, 7).cpu().numpy(),
                               d - np.log(np.exp(d).sum(-1, keepdims=True)), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(cl.entropy(torch.from_numpy(ref).cuda()).cpu().numpy(),
                               oracle.normed_entropy(ref), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('c,h,w', [(6, 48, 64), (13, 32, 80)])
def test_bayes_score_fused_tail_equals_generic_route(c, h, w):
    """BayesFusion.score() runs decode + decision table + confusion matrix as one kernel
    (xv_bayes_decode_score); the generic route (expert labels -> xv_bayes_fuse_lut ->
    xv_confusion_accumulate) must give the identical matrix, and both the oracle's count."""
    from xview.models import get_model
    from modular_semantic_segmentation_b200 import device as dev
    rng = np.random.default_rng(c)
    n = 3
    data = _data(rng, n, h, w, c)
    params = _trained_like(rng, c)
    cms = {m: rng.integers(0, 60, size=(c, c)).astype(np.float64) + 150 * np.eye(c)
           for m in ('rgb', 'depth')}
    common = dict(confusion_matrices=cms, data_description=_description(c),
                  prefixes={'rgb': 'rgb', 'depth': 'depth'}, expert_model='fcn', num_units=NU,
                  num_channels={'rgb': 3, 'depth': 1}, batchsize=2)
    with get_model('bayes_fusion')(**common) as net:
        _load(net, params)
        before = dev.launch_count()
        _, cm_fused = net.score(data)
        launches_fused = dev.launch_count() - before
        pred = net.predict(data)
        # the kernel can also return the fused labels it counted
        experts = [net._experts[m] for m in net.modalities]
        batch = net._to_device({k: v[:2] for k, v in data.items()})
        for m, e in zip(net.modalities, experts):
            e.forward(batch[m], want=())
        scratch = torch.zeros((c, c), dtype=torch.int64, device='cuda')
        fused = dev.bayes_decode_score(experts, net._lut, c, batch['labels'], scratch,
                                       want_fused=True)
        np.testing.assert_array_equal(fused.cpu().numpy(), pred[:2])
    with get_model('bayes_fusion')(fused_score_tail=False, **common) as net:
        _load(net, params)
        before = dev.launch_count()
        _, cm_generic = net.score(data)
        launches_generic = dev.launch_count() - before
    np.testing.assert_array_equal(cm_fused, cm_generic)
    np.testing.assert_array_equal(cm_fused, oracle.confusion_matrix(data['labels'], pred, c))
    assert launches_fused < launches_generic


@pytest.mark.parametrize('c,h,w', [(12, 48, 64), (5, 32, 80)])
def test_dirichlet_fused_tail_equals_generic_route(c, h, w):
    """DirichletFusion runs decode + softmax of both experts + Dirichlet fusion (+ confusion
    matrix) as one kernel (xv_dirichlet_decode_score); labels and matrices must be identical to
    the generic route (expert probabilities in HBM -> xv_dirichlet_fuse_exact ->
    xv_confusion_accumulate) and to the float32 oracle applied to the device's probabilities."""
    from xview.models import get_model
    from modular_semantic_segmentation_b200 import device as dev
    rng = np.random.default_rng(100 + c)
    n = 3
    data = _data(rng, n, h, w, c)
    params = _trained_like(rng, c)
    fit = {m: 1.0 + rng.gamma(2.0, 2.0, size=(c, c)) + 3 * np.eye(c) for m in ('rgb', 'depth')}
    fit['class_counts'] = rng.integers(10, 1000, size=c).astype(np.float64)
    common = dict(data_description=_description(c), modalities=['rgb', 'depth'], expert_model='fcn',
                  num_units=NU, num_channels={'rgb': 3, 'depth': 1}, batchsize=2,
                  dirichlet_params=fit)
    with get_model('dirichlet_mix')(**common) as net:
        _load(net, params)
        before = dev.launch_count()
        _, cm_fused = net.score(data)
        launches_fused = dev.launch_count() - before
        pred_fused = net.predict(data)
    with get_model('dirichlet_mix')(fused_score_tail=False, **common) as net:
        _load(net, params)
        before = dev.launch_count()
        _, cm_generic = net.score(data)
        launches_generic = dev.launch_count() - before
        pred_generic = net.predict(data)
        probs = [net._experts[m].forward(torch.from_numpy(data[m]).cuda(),
                                         want=('prob',))['prob'].cpu().numpy()
                 for m in ('rgb', 'depth')]
    np.testing.assert_array_equal(pred_fused, pred_generic)
    np.testing.assert_array_equal(cm_fused, cm_generic)
    np.testing.assert_array_equal(cm_fused, oracle.confusion_matrix(data['labels'], pred_fused, c))
    ref = oracle.dirichlet_fusion_f32(probs, [fit['rgb'], fit['depth']],
                                      oracle.dirichlet_prior(fit['class_counts']))
    np.testing.assert_array_equal(pred_fused, oracle.argmax_first(ref))
    assert launches_fused < launches_generic


def test_cuda_graph_score_step_equals_eager():
    """BaseModel.capture_score_step: the whole score() step of a device-resident batch replayed
    from one CUDA graph accumulates the same confusion matrix as the eager launches."""
    from xview.models import get_model
    c, n, h, w = 6, 4, 48, 64
    rng = np.random.default_rng(41)
    data = _data(rng, n, h, w, c)
    params = _trained_like(rng, c)
    cms = {m: rng.integers(0, 60, size=(c, c)).astype(np.float64) + 150 * np.eye(c)
           for m in ('rgb', 'depth')}
    with get_model('bayes_fusion')(
            confusion_matrices=cms, data_description=_description(c),
            prefixes={'rgb': 'rgb', 'depth': 'depth'}, expert_model='fcn', num_units=NU,
            num_channels={'rgb': 3, 'depth': 1}, batchsize=n) as net:
        _load(net, params)
        batch = net._to_device(data)
        eager = torch.zeros((c, c), dtype=torch.int64, device='cuda')
        net.score_batch_on_device(batch, eager)
        cm = torch.zeros((c, c), dtype=torch.int64, device='cuda')
        graph, kernels = net.capture_score_step(batch, cm)
        assert kernels == 33 and int(cm.sum()) == 0          # capture leaves the matrix empty
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        np.testing.assert_array_equal(cm.cpu().numpy(), 3 * eager.cpu().numpy())


@pytest.mark.parametrize('model', ['bayes_fusion', 'dirichlet_mix', 'average_fusion',
                                   'variance_fusion'])
def test_concurrent_experts_equal_sequential_experts(model):
    """Small batches run the experts of the modalities on separate CUDA streams (forked from and
    joined to the caller's stream, `overlap_experts`); results are those of the sequential
    schedule bit for bit, through predict(), score() and repeated calls."""
    from xview.models import get_model
    c, n, h, w = 6, 3, 48, 64
    rng = np.random.default_rng(5)
    data = _data(rng, n, h, w, c)
    params = _trained_like(rng, c)
    extra = {}
    if model == 'bayes_fusion':
        extra['confusion_matrices'] = {m: rng.integers(0, 60, size=(c, c)).astype(np.float64) +
                                       150 * np.eye(c) for m in ('rgb', 'depth')}
    if model == 'dirichlet_mix':
        extra['dirichlet_params'] = {m: 1.0 + rng.gamma(2.0, 2.0, size=(c, c)) + 4.0 * np.eye(c)
                                     for m in ('rgb', 'depth')}
        extra['dirichlet_params']['class_counts'] = rng.integers(100, 1000, size=c).astype(np.float64)
        extra['modalities'] = ['rgb', 'depth']
    else:
        extra['prefixes'] = {'rgb': 'rgb', 'depth': 'depth'}
    if model == 'variance_fusion':
        extra.update(num_samples=5, dropout_rate=0.4, deterministic_dropout=True, seed=3)
    results = {}
    for overlap in (False, True):
        with get_model(model)(data_description=_description(c), expert_model='fcn', num_units=NU,
                              num_channels={'rgb': 3, 'depth': 1}, batchsize=2,
                              overlap_experts=overlap, **extra) as net:
            _load(net, params)
            preds = [net.predict({'rgb': data['rgb'], 'depth': data['depth']}) for _ in range(3)]
            cms = [net.score(data)[1] for _ in range(2)]
        for p in preds[1:]:
            np.testing.assert_array_equal(p, preds[0])
        np.testing.assert_array_equal(cms[0], cms[1])
        results[overlap] = (preds[0], cms[0])
    np.testing.assert_array_equal(results[True][0], results[False][0])
    np.testing.assert_array_equal(results[True][1], results[False][1])
    assert results[True][1].sum() == (data['labels'] >= 0).sum()


def test_bayes_fusion_insight_dump(tmp_path):
    """experiments/bayes_fusion.py:47-69 (`collect_data`): get_insight returns the experts'
    probabilities, the log-likelihood rows and conditionals bayes_fusion selects for their
    decisions (bit-equal to the oracle's float32 rule on the same decisions) and the fused
    prediction; records.dump_bayes_insight stores them the way the command does."""
    from xview.models import get_model
    from modular_semantic_segmentation_b200 import records
    c, n, h, w = 6, 3, 32, 48
    rng = np.random.default_rng(12)
    data = _data(rng, n, h, w, c)
    params = _trained_like(rng, c)
    cms = {m: rng.integers(0, 60, size=(c, c)).astype(np.float64) + 150 * np.eye(c)
           for m in ('rgb', 'depth')}
    with get_model('bayes_fusion')(
            confusion_matrices=cms, data_description=_description(c),
            prefixes={'rgb': 'rgb', 'depth': 'depth'}, expert_model='fcn', num_units=NU,
            num_channels={'rgb': 3, 'depth': 1}, batchsize=2) as net:
        _load(net, params)
        probs, likelihoods, conditionals, prediction = net.get_insight(data)
        fused = net.predict(data)
        batches = [{k: v[i:i + 2] for k, v in data.items()} for i in (0, 2)]
        paths = records.dump_bayes_insight(net, batches, str(tmp_path / 'insight'))
    assert probs.shape == likelihoods.shape == conditionals.shape == (2, n, h, w, c)
    np.testing.assert_allclose(probs.sum(-1), 1.0, atol=1e-5)
    np.testing.assert_array_equal(prediction, fused)
    decisions = [probs[i].argmax(-1) for i in range(2)]
    tables = [cms[m].astype('float32').T for m in ('rgb', 'depth')]
    ref_score, ref_ll, ref_cond = oracle.bayes_fusion(decisions, tables, 'data')
    for i in range(2):
        np.testing.assert_array_equal(likelihoods[i], ref_ll[i].astype(np.float32))
        np.testing.assert_array_equal(conditionals[i], ref_cond[i].astype(np.float32))
    np.testing.assert_array_equal(prediction, oracle.argmax_first(ref_score))
    assert [os.path.basename(p) for p in paths] == ['predictions.npz', 'likelihoods.npz',
                                                    'conditionals.npz', 'probs.npz']
    stored = np.load(paths[0])
    assert sorted(stored.files) == ['arr_0', 'arr_1']
    np.testing.assert_array_equal(np.concatenate([stored['arr_0'], stored['arr_1']]), fused)
    assert np.load(paths[3])['arr_1'].shape == (2, 1, h, w, c)


def test_bayesian_fcn_sampling_uncertainty():
    """bayesian_fcn.py:9-57: mean and uncertainty measures of the MC-dropout samples against the
    oracle's restatement evaluated on the very same samples (fixed Philox seed), through the
    function and through the model class."""
    from xview.models import get_model
    from xview.models.bayesian_fcn import sampling_uncertainty
    c, n, h, w, t = 6, 2, 32, 48, 7
    rng = np.random.default_rng(21)
    data = _data(rng, n, h, w, c)
    params = {k: v for k, v in _trained_like(rng, c).items() if k.startswith('rgb/')}
    layers = ['pool3', 'pool4', 'conv5_3', 'features']
    with get_model('bayesian_fcn')('rgb', _description(c), 'rgb', num_units=NU, num_samples=t,
                                   dropout_rate=0.3, dropout_layers=layers, batchsize=n,
                                   deterministic_dropout=True, seed=2) as net:
        _load(net, params)
        expert = net._experts['rgb']
        x = torch.from_numpy(data['rgb']).cuda()
        from xview.models.variance_mix import mc_dropout_seed
        seed = mc_dropout_seed(net, 0)
        samples = expert.forward(x, want=('prob',), dropout={
            'rate': 0.3, 'layers': ['pool3', 'conv5_3', 'features'], 'num_samples': t,
            'seed': seed})['prob'].cpu().numpy().reshape(t, n, h, w, c)
        mean, unc = sampling_uncertainty(x, expert, t, c, dropout_rate=0.3, dropout_layers=layers,
                                         seed=seed)
        mean, unc = mean.cpu().numpy(), {k: v.cpu().numpy() for k, v in unc.items()}
        pred = net.predict({'rgb': data['rgb']})
        entropy = net.predict({'rgb': data['rgb']}, output_attr='entropy')
        measures, cm = net.score(data)
    assert samples.shape == (t, n, h, w, c) and np.abs(samples[0] - samples[1]).max() > 1e-4
    ref_mean, ref_unc = oracle.sampling_uncertainty(samples.astype(np.float64))
    np.testing.assert_allclose(mean, ref_mean, atol=1e-6)
    for key in ('entropy', 'cond_entropy', 'variance'):
        np.testing.assert_allclose(unc[key], ref_unc[key], atol=2e-6, err_msg=key)
    np.testing.assert_array_equal(pred, mean.argmax(-1))
    np.testing.assert_array_equal(entropy, unc['entropy'])
    np.testing.assert_array_equal(cm, oracle.confusion_matrix(data['labels'], pred, c))
