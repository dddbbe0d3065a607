"""GPU parity: the Adapnet expert (xv_adapnet_create / xv_fcn_forward) against oracle/adapnet.py."""
import numpy as np
import pytest
import torch

import oracle
from util import cuda

pytestmark = pytest.mark.gpu

LAYERS = (['block_0_1', 'block_0_2', 'block_0_pool'] + ['block_%d' % i for i in range(1, 17)] +
          ['shortcut', 'merge'])


@pytest.fixture(scope='module')
def dev():
    from modular_semantic_segmentation_b200 import device
    device.init()
    return device


def _net(dev, precision, cin, nu, c, rng):
    # gain > 1 keeps the activations O(1) through 50 layers of ReLU
    params = oracle.adapnet_params('m', cin, nu, c, rng, gain=1.3)
    net = dev.FcnExpert(cin, nu, c, precision=precision, arch='adapnet')
    net.set_params({k.split('/', 1)[1]: v for k, v in params.items()})
    return net, params


@pytest.mark.parametrize('cin,nu,c,h,w', [(3, 20, 14, 32, 48), (1, 64, 12, 48, 32)])
def test_adapnet_fp32_validation_mode_matches_oracle(dev, cin, nu, c, h, w):
    rng = np.random.default_rng(cin)
    net, params = _net(dev, 'fp32', cin, nu, c, rng)
    x = rng.uniform(0, 1, size=(2, h, w, cin)).astype(np.float32)
    ref = oracle.adapnet(x, params, 'm', nu, c)
    out = net.forward(cuda(x), want=('score', 'prob', 'label'))
    for name in LAYERS:
        got = net.layer(name)
        scale = max(np.abs(ref[name]).max(), 1e-6)
        np.testing.assert_allclose(got, ref[name], rtol=0, atol=2e-4 * scale, err_msg=name)
    score = out['score'].cpu().numpy()
    np.testing.assert_allclose(score, ref['score'], rtol=0,
                               atol=2e-4 * np.abs(ref['score']).max())
    np.testing.assert_allclose(out['prob'].cpu().numpy(), oracle.softmax(ref['score']), rtol=0,
                               atol=1e-4)
    assert (out['label'].cpu().numpy() == ref['score'].argmax(-1)).mean() > 0.999


@pytest.mark.parametrize('cin,nu,c,h,w', [(3, 20, 14, 64, 96), (1, 64, 12, 96, 64),
                                          (3, 64, 12, 128, 160)])
def test_adapnet_bf16_tcgen05_matches_oracle(dev, cin, nu, c, h, w):
    """bf16 tensor-core path: every block output close to the fp32 oracle relative to the layer's
    own scale, probabilities within the 2e-2 bf16 tolerance."""
    rng = np.random.default_rng(20 + cin + nu)
    net, params = _net(dev, 'bf16', cin, nu, c, rng)
    hi = 255.0 if cin == 3 else 65535.0
    x = rng.integers(0, int(hi) + 1, size=(2, h, w, cin)).astype(np.float32)
    params['m/block_0_1/kernel'] = params['m/block_0_1/kernel'] / np.float32(hi)
    net.set_param('block_0_1/kernel', params['m/block_0_1/kernel'])
    ref = oracle.adapnet(x, params, 'm', nu, c)
    # "trained-like" logits: the random init gives class scores of arbitrary scale; the batch norm
    # of the last layer is rescaled so that |score| <= 3 and the plain north_star tolerance
    # (2e-2 abs on the probabilities) is the meaningful bar
    f = np.float32(min(1.0, 3.0 / np.abs(ref['score']).max()))
    for leaf in ('gamma', 'beta'):
        name = 'second_deconvolution_upconv/' + leaf
        params['m/' + name] = params['m/' + name] * f
        net.set_param(name, params['m/' + name])
    ref = oracle.adapnet(x, params, 'm', nu, c)
    out = net.forward(cuda(x), want=('score', 'prob', 'label'))
    for name in LAYERS:
        got = net.layer(name)
        assert got.shape == ref[name].shape, name
        err = np.abs(got - ref[name])
        scale = np.abs(ref[name]).max()
        # bf16 storage between ~55 layers: mean error stays a small fraction of the layer scale
        assert err.mean() < 6e-3 * scale and err.max() < 8e-2 * scale, (
            name, float(err.mean() / scale), float(err.max() / scale))
    prob_ref = oracle.softmax(ref['score'])
    prob = out['prob'].cpu().numpy()
    tol = 2e-2
    assert np.abs(prob - prob_ref).max() < tol, float(np.abs(prob - prob_ref).max())
    label = out['label'].cpu().numpy()
    margin = np.sort(prob_ref, -1)
    decisive = (margin[..., -1] - margin[..., -2]) > 2 * tol
    assert (label == prob_ref.argmax(-1))[decisive].all()
    assert (label == prob_ref.argmax(-1)).mean() > 0.97


def test_adapnet_model_class_and_fusion_expert():
    """Adapnet.predict / score and BayesFusion(expert_model='adapnet') against the oracle."""
    from xview.models import get_model
    c, nu, h, w = 6, 16, 64, 64
    rng = np.random.default_rng(3)
    description = [{'rgb': None, 'depth': None, 'labels': None},
                   {'rgb': [None, None, 3], 'depth': [None, None, 1], 'labels': [None, None]}, c]
    params = {}
    for prefix, cin in (('rgb', 3), ('depth', 1)):
        params.update(oracle.adapnet_params(prefix, cin, nu, c, rng, gain=1.3))
    data = {'rgb': rng.uniform(0, 1, size=(3, h, w, 3)).astype(np.float32),
            'depth': rng.uniform(0, 1, size=(3, h, w, 1)).astype(np.float32),
            'labels': rng.integers(0, c, size=(3, h, w)).astype(np.int32)}
    ref = {m: oracle.adapnet(data[m], params, m, nu, c)['score'] for m in ('rgb', 'depth')}
    with get_model('adapnet')(data_description=description, modality='rgb', num_units=nu,
                              precision='fp32') as net:
        net.variables.update({k: v for k, v in params.items() if k.startswith('rgb/')})
        net._push_variables()
        pred = net.predict(data)
        assert pred.dtype == np.int64 and pred.shape == (3, h, w)
        assert (pred == ref['rgb'].argmax(-1)).mean() > 0.999
        with pytest.raises(UserWarning):
            net.fit(data, 1)
    cms = {m: rng.integers(1, 50, size=(c, c)).astype(np.float64) + 100 * np.eye(c)
           for m in ('rgb', 'depth')}
    with get_model('bayes_fusion')(
            confusion_matrices=cms, data_description=description,
            prefixes={'rgb': 'rgb', 'depth': 'depth'}, expert_model='adapnet', num_units=nu,
            num_channels={'rgb': 3, 'depth': 1}, batchsize=2, precision='fp32') as net:
        net.variables.update(params)
        net._push_variables()
        fused = net.predict(data)
        measures, cm = net.score(data)
    labels = [ref[m].argmax(-1) for m in ('rgb', 'depth')]
    tables = [cms[m].astype('float32').T for m in ('rgb', 'depth')]    # bayes_mix.py:138-140
    want = oracle.argmax_first(oracle.bayes_fusion(labels, tables)[0])
    assert (fused == want).mean() > 0.999
    np.testing.assert_array_equal(cm, oracle.confusion_matrix(data['labels'], fused, c))
