"""GPU checks at the full sizes of BASELINE.json's configs (768x384 frames, num_units 64, 12
classes) through size-independent properties: the oracle would need minutes per frame there, so
these tests assert what must hold at any size - invariance to how a data set is split into batches
and uploads, additivity of the confusion matrix, its checksum, permutation equivariance, and
agreement of the fused labels with the oracle's fusion rule applied to the device's own expert
outputs."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

C, NU = 12, 64


def _description():
    return ({'rgb': np.float32, 'depth': np.float32, 'labels': np.int32},
            {'rgb': (None, None, 3), 'depth': (None, None, 1), 'labels': (None, None)}, C)


def _data(rng, n, h, w):
    return {'rgb': rng.integers(0, 256, size=(n, h, w, 3)).astype(np.float32),
            'depth': rng.integers(0, 65536, size=(n, h, w, 1)).astype(np.float32),
            'labels': rng.integers(-1, C, size=(n, h, w)).astype(np.int32)}


def _cms(rng):
    return {m: rng.integers(0, 60, size=(C, C)).astype(np.float64) + 400 * np.eye(C)
            for m in ('rgb', 'depth')}


def _bayes(cms, batchsize, **extra):
    from xview.models import get_model
    return get_model('bayes_fusion')(
        confusion_matrices=cms, data_description=_description(),
        prefixes={'rgb': 'rgb', 'depth': 'depth'}, expert_model='fcn', num_units=NU,
        num_channels={'rgb': 3, 'depth': 1}, batchsize=batchsize, seed=11, **extra)


@pytest.mark.parametrize('h,w', [(768, 384), (384, 768)])
def test_bayes_fusion_full_size_properties(h, w):
    """configs[1] at both frame orientations, 6 frames."""
    rng = np.random.default_rng(h)
    n = 6
    data = _data(rng, n, h, w)
    cms = _cms(rng)
    with _bayes(cms, 4) as net:               # batches of 4 + 2 (ragged), default upload pieces
        fused = net.predict(data)
        measures, cm = net.score(data)
        experts = {m: net.expert_outputs[m]['classification'].cpu().numpy().astype(np.int64)
                   for m in net.modalities}   # last batch = frames 4, 5
    with _bayes(cms, 1, upload_split=1) as net:       # frame by frame, one upload each
        fused_1 = net.predict(data)
        _, cm_a = net.score({k: v[:2] for k, v in data.items()})
        _, cm_b = net.score({k: v[2:] for k, v in data.items()})
        order = np.array([3, 0, 5, 1, 4, 2])
        fused_perm = net.predict({k: v[order] for k, v in data.items()})
    # the same frames give the same labels however they are batched, split for upload or ordered
    np.testing.assert_array_equal(fused, fused_1)
    np.testing.assert_array_equal(fused_perm, fused[order])
    # confusion matrix: additive over a partition of the data, rows = labels, checksum = number
    # of labelled pixels, and identical to the host count of the returned labels
    np.testing.assert_array_equal(cm, cm_a + cm_b)
    assert cm.sum() == (data['labels'] >= 0).sum()
    np.testing.assert_array_equal(cm, oracle.confusion_matrix(data['labels'], fused, C))
    np.testing.assert_array_equal(cm.sum(1), np.bincount(data['labels'][data['labels'] >= 0],
                                                         minlength=C))
    ref_measures = oracle.score_measures(cm)
    assert measures['mean_IoU'] == ref_measures['mean_IoU']
    # fusion rule on the device's own expert labels: bit-exact against the oracle
    tables = [cms[m].astype('float32').T for m in ('rgb', 'depth')]
    want = oracle.argmax_first(oracle.bayes_fusion([experts['rgb'], experts['depth']], tables)[0])
    np.testing.assert_array_equal(fused[4:], want)


def test_kernel_variants_agree_at_full_size():
    """One 768x384 rgb frame through the default kernels (CTA pairs, halo operands, TMA-staged
    conv1_1) and through the previous variants: probabilities within the bf16 tolerance."""
    from modular_semantic_segmentation_b200 import device as dev
    from modular_semantic_segmentation_b200.models.simple_fcn import build_expert
    rng = np.random.default_rng(5)
    expert, variables = build_expert('rgb', 3, NU, C, rng=rng)
    variables['rgb/conv1_1/kernel'] = variables['rgb/conv1_1/kernel'] / np.float32(255.0)
    expert.set_params({k[4:]: v * (1.45 if k.endswith('kernel') and 'up' not in k else 1.0)
                       for k, v in variables.items()})
    x = torch.from_numpy(rng.integers(0, 256, size=(1, 768, 384, 3)).astype(np.float32)).cuda()
    base = expert.forward(x, want=('prob', 'label'))
    for flags in (32 | 64, 128):
        dev.set_debug_flags(flags)
        alt = expert.forward(x, want=('prob', 'label'))
        dev.set_debug_flags(0)
        assert (alt['prob'] - base['prob']).abs().max().item() < 2e-2
        assert (alt['label'] == base['label']).float().mean().item() > 0.99
    np.testing.assert_allclose(base['prob'].sum(-1).cpu().numpy(), 1.0, atol=1e-5)
    expert.close()


def test_mc_dropout_full_size_properties():
    """configs[2] shape: T = 20 MC-dropout samples of a 768x384 frame; the Monte-Carlo moments
    obey their definitions at any size, and variance fusion returns valid label maps."""
    from xview.models import get_model
    from modular_semantic_segmentation_b200.models.simple_fcn import build_expert
    rng = np.random.default_rng(9)
    data = _data(rng, 2, 768, 384)
    expert, variables = build_expert('depth', 1, NU, C, rng=rng)
    variables['depth/conv1_1/kernel'] = variables['depth/conv1_1/kernel'] / np.float32(65535.0)
    expert.set_params({k[6:]: v for k, v in variables.items()})
    x = torch.from_numpy(data['depth'][:1]).cuda()
    drop = {'rate': 0.5, 'layers': ['pool3'], 'num_samples': 20, 'seed': 4}
    out = expert.forward(x, want=('prob', 'mean_prob', 'var_prob', 'mean_var'), dropout=drop)
    samples = out['prob'].double()                       # [20, H, W, C]
    assert samples.shape == (20, 768, 384, C)
    mean, var = samples.mean(0), samples.var(0, unbiased=False)
    assert (out['mean_prob'][0].double() - mean).abs().max().item() < 1e-5
    assert (out['var_prob'][0].double() - var).abs().max().item() < 1e-5
    assert (out['mean_var'][0].double() - var.mean(-1)).abs().max().item() < 1e-5
    assert (out['mean_prob'].sum(-1) - 1).abs().max().item() < 1e-4
    again = expert.forward(x, want=('mean_var',), dropout=drop)      # same seed -> same masks
    assert torch.equal(again['mean_var'], out['mean_var'])
    expert.close()
    with get_model('variance_fusion')(
            data_description=_description(), prefixes={'rgb': 'rgb', 'depth': 'depth'},
            expert_model='fcn', num_units=NU, num_channels={'rgb': 3, 'depth': 1}, batchsize=1,
            num_samples=20, dropout_rate=0.5, seed=3) as net:
        pred = net.predict(data)
        measures, cm = net.score(data)
    assert pred.shape == (2, 768, 384) and pred.dtype == np.int64
    assert pred.min() >= 0 and pred.max() < C
    np.testing.assert_array_equal(cm, oracle.confusion_matrix(data['labels'], pred, C))
