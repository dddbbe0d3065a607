"""GPU parity: the whole FCN expert (xv_fcn_forward) against the oracle, layer by layer."""
import numpy as np
import pytest
import torch

import oracle
from util import cuda

pytestmark = pytest.mark.gpu

NU, C = 8, 5
LAYERS = ['conv1_1', 'conv1_2', 'pool1', 'conv2_2', 'pool2', 'conv3_3', 'pool3', 'conv4_3',
          'pool4', 'conv5_3', 'score_conv4', 'score_conv5', 'fused']


def _net(dev, precision, cin, rng, batchnorm=False, gain=1.45):
    params = oracle.glorot_fcn_params('m', cin, NU, C, rng, gain=gain, bias_scale=0.05,
                                      batchnorm=batchnorm)
    net = dev.FcnExpert(cin, NU, C, batchnorm=batchnorm, precision=precision)
    net.set_params({k.split('/', 1)[1]: v for k, v in params.items()})
    return net, params


@pytest.fixture(scope='module')
def dev():
    from modular_semantic_segmentation_b200 import device
    device.init()
    return device


@pytest.mark.parametrize('cin,h,w', [(3, 32, 48), (1, 48, 32)])
def test_fcn_fp32_validation_mode_matches_oracle(dev, cin, h, w):
    """Probabilities within 1e-4 abs of the oracle (north_star fp32 validation tolerance)."""
    rng = np.random.default_rng(cin)
    net, params = _net(dev, 'fp32', cin, rng)
    x = rng.uniform(0, 1, size=(2, h, w, cin)).astype(np.float32)
    ref = oracle.test_pipeline(x, params, 'm', NU, C)
    out = net.forward(cuda(x), want=('score', 'prob', 'label'))
    for name in LAYERS + ['upscore', 'score']:
        got = net.layer(name)
        scale = max(np.abs(ref[name]).max(), 1e-6)
        np.testing.assert_allclose(got, ref[name], rtol=0, atol=1e-4 * scale, err_msg=name)
    np.testing.assert_allclose(out['prob'].cpu().numpy(), ref['prob'], rtol=0, atol=1e-4)
    agree = (out['label'].cpu().numpy() == ref['classification']).mean()
    assert agree > 0.999


@pytest.mark.parametrize('fused_pool', [True, False])
@pytest.mark.parametrize('cin,h,w', [(3, 64, 96), (1, 96, 64), (3, 48, 80)])
def test_fcn_bf16_tcgen05_matches_oracle(dev, cin, h, w, fused_pool):
    """Probabilities within 2e-2 abs of the fp32 oracle (north_star bf16 tolerance).  With
    pooling fused into the conv epilogues conv1_2 / conv2_2 / conv3_3 are never materialised
    (conv4_3 is: it also feeds score_conv4)."""
    dev.set_debug_flags(0 if fused_pool else 1)
    rng = np.random.default_rng(10 + cin)
    net, params = _net(dev, 'bf16', cin, rng)
    hi = 255.0 if cin == 3 else 65535.0
    x = rng.integers(0, int(hi) + 1, size=(2, h, w, cin)).astype(np.float32)
    # fold the input range into conv1_1 so activations are O(1) ("trained-like" variant)
    params['m/conv1_1/kernel'] = params['m/conv1_1/kernel'] / np.float32(hi)
    net.set_param('conv1_1/kernel', params['m/conv1_1/kernel'])
    ref = oracle.test_pipeline(x, params, 'm', NU, C)
    out = net.forward(cuda(x), want=('score', 'prob', 'label'))
    report = []
    for name in LAYERS:
        if fused_pool and name in ('conv1_2', 'conv2_2', 'conv3_3'):
            with pytest.raises(Exception):
                net.layer(name)
            continue
        got = net.layer(name)
        assert got.shape == ref[name].shape, name
        err = np.abs(got - ref[name]).max() / max(np.abs(ref[name]).max(), 1e-6)
        report.append('%s %.4f' % (name, err))
        assert err < 0.05, (name, report)
    print('relative max-abs error per layer:', ', '.join(report))
    np.testing.assert_allclose(out['prob'].cpu().numpy(), ref['prob'], rtol=0, atol=2e-2)
    np.testing.assert_allclose(out['prob'].cpu().numpy().sum(-1), 1.0, atol=1e-5)
    agree = (out['label'].cpu().numpy() == ref['classification']).mean()
    assert agree > 0.97
    # uint8 labels carry the same decisions
    out8 = net.forward(cuda(x), want=('label',), label_dtype=torch.uint8)
    np.testing.assert_array_equal(out8['label'].cpu().numpy(), out['label'].cpu().numpy())
    dev.set_debug_flags(0)


@pytest.mark.parametrize('cin', [3, 1])
def test_fcn_kernel_variants_agree(dev, cin):
    """The previous kernel variants (debug bit5: conv1_1 packers reading global memory, bit6:
    nine shifted tiles instead of three patch copies, bit7: single-CTA instead of cta_group::2
    MMAs) compute the same network up to summation order and bf16 rounding."""
    rng = np.random.default_rng(40 + cin)
    net, params = _net(dev, 'bf16', cin, rng)
    hi = 255.0 if cin == 3 else 65535.0
    x = cuda(rng.integers(0, int(hi) + 1, size=(2, 48, 80, cin)).astype(np.float32))
    net.set_param('conv1_1/kernel', params['m/conv1_1/kernel'] / np.float32(hi))
    base = net.forward(x, want=('prob', 'label'))
    base_c11 = net.layer('conv1_1')
    for flags in (32, 64, 96, 128):     # bit7: single-CTA kernel instead of the CTA-pair one
        dev.set_debug_flags(flags)
        alt = net.forward(x, want=('prob', 'label'))
        alt_c11 = net.layer('conv1_1')
        dev.set_debug_flags(0)
        # conv1_1: the bias enters as hi + lo bf16 inside the MMA vs an fp32 add
        np.testing.assert_allclose(alt_c11, base_c11, rtol=0, atol=2.0 ** -7 * np.abs(base_c11).max())
        np.testing.assert_allclose(alt['prob'].cpu().numpy(), base['prob'].cpu().numpy(), rtol=0,
                                   atol=1e-2)
        assert (alt['label'] == base['label']).float().mean().item() > 0.99


@pytest.mark.parametrize('n,h,w', [(2, 64, 48), (1, 48, 32), (3, 80, 16), (1, 128, 16)])
def test_conv1_2_row_pair_kernel_matches_the_half_empty_one(dev, n, h, w):
    """conv1_2 + pool1 through the row-pair kernel (accumulator lanes 64..127 carry the next
    output row) against the transposed-role kernel it replaces (debug bit10) and against an fp64
    convolution of the same bf16 operands: same products, another summation order - at most one
    bf16 ulp apart, and only rarely.  Heights that are not a multiple of the 32-row tile included."""
    rng = np.random.default_rng(h + w)
    net, params = _net(dev, 'bf16', 3, rng)
    x = cuda(rng.uniform(0, 1, size=(n, h, w, 3)).astype(np.float32))
    net.forward(x, want=('label',))
    new, c11 = net.layer('pool1'), net.layer('conv1_1')
    dev.set_debug_flags(1024)
    net.forward(x, want=('label',))
    old = net.layer('pool1')
    dev.set_debug_flags(0)
    assert new.shape == old.shape == (n, h // 2, w // 2, 64)
    ulp = 2.0 ** -7 * np.maximum(np.abs(old), 2.0 ** -126)
    assert (np.abs(new - old) <= ulp).all()
    assert (new == old).mean() > 0.98
    # exact reference from the stored bf16 conv1_1 activation and bf16-rounded weights
    wq = torch.from_numpy(params['m/conv1_2/kernel']).bfloat16().double()      # HWIO
    ref = torch.nn.functional.conv2d(torch.from_numpy(c11).double().permute(0, 3, 1, 2),
                                     wq.permute(3, 2, 0, 1), padding=1)
    ref = torch.relu(ref + torch.from_numpy(params['m/conv1_2/bias']).double()[None, :, None, None])
    ref = torch.nn.functional.max_pool2d(ref, 2).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(new, ref, rtol=2.0 ** -7, atol=1e-6)


def test_fcn_batchnorm_fp32(dev):
    rng = np.random.default_rng(5)
    net, params = _net(dev, 'fp32', 3, rng, batchnorm=True)
    x = rng.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    ref = oracle.fcn(x, params, 'm', NU, C, batchnorm=True)
    out = net.forward(cuda(x), want=('score',))
    np.testing.assert_allclose(out['score'].cpu().numpy(), ref['score'], rtol=0,
                               atol=1e-4 * np.abs(ref['score']).max())


def test_fcn_batchnorm_bf16_folds_into_tensor_core_convs(dev):
    rng = np.random.default_rng(6)
    net, params = _net(dev, 'bf16', 3, rng, batchnorm=True)
    x = rng.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    ref = oracle.fcn(x, params, 'm', NU, C, batchnorm=True)
    out = net.forward(cuda(x), want=('prob',))
    np.testing.assert_allclose(out['prob'].cpu().numpy(), oracle.softmax(ref['score']), rtol=0,
                               atol=2e-2)


def _masks(rng, t, n, h, w, rate, sites):
    shapes = {'pool3': (h // 8, w // 8, 256), 'pool4': (h // 16, w // 16, 512),
              'conv4_3': (h // 8, w // 8, 512), 'conv5_3': (h // 16, w // 16, 512),
              'features': (h // 8, w // 8, NU)}
    need = set(sites)
    if 'pool3' in need:
        need.add('pool4')
    return {s: (rng.random((t * n,) + shapes[s]) >= rate).astype(np.uint8) for s in need}


@pytest.mark.parametrize('precision,sites', [('fp32', ['pool3']), ('fp32', ['conv4_3', 'features']),
                                             ('bf16', ['pool3']), ('bf16', ['conv5_3'])])
def test_mc_dropout_with_shared_masks(dev, precision, sites):
    """Identical external keep-masks -> every MC sample equals the oracle's dropout pass, and
    the fused moments equal the moments of those samples."""
    rng = np.random.default_rng(len(sites) * 7 + (precision == 'bf16'))
    t, n, h, w, rate = 3, 2, 32, 48, 0.3
    net, params = _net(dev, precision, 3, rng)
    x = rng.uniform(0, 1, size=(n, h, w, 3)).astype(np.float32)
    masks = _masks(rng, t, n, h, w, rate, sites)
    out = net.forward(cuda(x), want=('prob', 'mean_prob', 'var_prob', 'mean_var'),
                      dropout={'rate': rate, 'layers': sites, 'num_samples': t, 'masks': masks})
    got = out['prob'].cpu().numpy().reshape(t, n, h, w, C)
    ref = []
    for i in range(t):
        m = {s: v[i * n:(i + 1) * n] for s, v in masks.items()}
        ref.append(oracle.test_pipeline(x, params, 'm', NU, C, dropout_rate=rate,
                                        dropout_layers=sites, masks=m)['prob'])
    ref = np.stack(ref)
    tol = 1e-4 if precision == 'fp32' else 2e-2
    np.testing.assert_allclose(got, ref, rtol=0, atol=tol)
    mean_ref, var_ref = oracle.mc_moments(got.astype(np.float64), 0)
    np.testing.assert_allclose(out['mean_prob'].cpu().numpy(), mean_ref, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out['var_prob'].cpu().numpy(), var_ref, rtol=1e-3, atol=1e-7)
    np.testing.assert_allclose(out['mean_var'].cpu().numpy(), var_ref.mean(-1), rtol=1e-3,
                               atol=1e-7)


@pytest.mark.parametrize('precision', ['bf16', 'fp32'])
def test_mc_dropout_with_dropout_free_leading_sample(dev, precision):
    """XV_DROP_FLAG_KEEP_FIRST: the dropout-free pass of variance_mix.py:68-69 rides along as a
    leading sample.  Its outputs equal a plain forward call bit for bit and the moments equal
    those of a separate MC call with the same external masks."""
    rng = np.random.default_rng(77)
    t, n, h, w, rate = 3, 2, 32, 48, 0.3
    net, params = _net(dev, precision, 3, rng)
    x = cuda(rng.uniform(0, 1, size=(n, h, w, 3)).astype(np.float32))
    masks = _masks(rng, t, n, h, w, rate, ['pool3'])
    cfg = {'rate': rate, 'layers': ['pool3'], 'num_samples': t, 'masks': masks}
    plain = net.forward(x, want=('prob', 'label', 'score'))
    mc = net.forward(x, want=('mean_prob', 'var_prob', 'mean_var'), dropout=cfg)
    both = net.forward(x, want=('prob', 'label', 'score', 'mean_prob', 'var_prob', 'mean_var'),
                       dropout=dict(cfg, with_deterministic=True))
    assert both['prob'].shape == (n, h, w, C)
    for key in ('prob', 'label', 'score'):
        assert torch.equal(both[key], plain[key]), key
    for key in ('mean_prob', 'var_prob', 'mean_var'):
        assert torch.equal(both[key], mc[key]), key
    # Philox masks: the leading sample stays dropout-free
    philox = net.forward(x, want=('prob',), dropout={'rate': rate, 'layers': ['pool3'],
                                                     'num_samples': t, 'seed': 3,
                                                     'with_deterministic': True})
    assert torch.equal(philox['prob'], plain['prob'])


def test_vector_dropout_and_staged_mc_decode_match_their_scalar_variants(dev):
    """The 8-elements-per-thread dropout kernel draws the same Philox / external masks as the
    scalar kernel (debug flag 256), bit for bit, also behind a dropout-free leading sample; the
    MC decode that stages eight samples per barrier (ex2.approx softmax) stays within 1e-6 of
    the per-sample-barrier kernel (flag 512) and does not depend on the outputs asked for."""
    rng = np.random.default_rng(31)
    t, n, h, w, rate = 11, 2, 32, 48, 0.5
    net, _ = _net(dev, 'bf16', 3, rng)
    x = cuda(rng.uniform(0, 1, size=(n, h, w, 3)).astype(np.float32))
    want = ('prob', 'mean_prob', 'var_prob', 'mean_var')
    for cfg in ({'rate': rate, 'layers': ['pool3', 'conv5_3'], 'num_samples': t, 'seed': 5},
                {'rate': rate, 'layers': ['pool3'], 'num_samples': t, 'seed': 6,
                 'with_deterministic': True},
                {'rate': rate, 'layers': ['conv4_3'], 'num_samples': t,
                 'masks': _masks(rng, t, n, h, w, rate, ['conv4_3'])}):
        new = {k: v.clone() for k, v in net.forward(x, want=want, dropout=cfg).items()}
        only = net.forward(x, want=('mean_var',), dropout=cfg)['mean_var'].clone()
        assert torch.equal(only, new['mean_var'])
        dev.set_debug_flags(256)
        scalar = {k: v.clone() for k, v in net.forward(x, want=want, dropout=cfg).items()}
        dev.set_debug_flags(512)
        slow = {k: v.clone() for k, v in net.forward(x, want=want, dropout=cfg).items()}
        dev.set_debug_flags(0)
        for key in want:
            assert torch.equal(scalar[key], new[key]), key
            assert (slow[key] - new[key]).abs().max().item() < 1e-6, key
        assert torch.equal(slow['prob'], new['prob'])


def test_fused_philox_dropout_statistics(dev):
    """Without external masks the fused Philox generator must drop ~rate of the units,
    independently per sample, and be reproducible for a fixed seed."""
    rng = np.random.default_rng(2)
    t, n, h, w, rate = 8, 1, 32, 32, 0.4
    net, _ = _net(dev, 'bf16', 3, rng)
    x = rng.uniform(0, 1, size=(n, h, w, 3)).astype(np.float32)
    cfg = {'rate': rate, 'layers': ['pool3'], 'num_samples': t, 'seed': 1234}
    a = net.forward(cuda(x), want=('prob',), dropout=cfg)['prob'].cpu().numpy()
    drop = net.layer('pool3_drop')
    src = net.layer('pool3')
    alive = src > 0
    kept = (drop.reshape((t,) + src.shape)[:, alive] > 0).mean()
    assert abs(kept - (1 - rate)) < 0.02, kept
    per_sample = drop.reshape(t, -1)
    assert not np.array_equal(per_sample[0], per_sample[1])
    b = net.forward(cuda(x), want=('prob',), dropout=cfg)['prob'].cpu().numpy()
    np.testing.assert_array_equal(a, b)
    cfg['seed'] = 99
    c = net.forward(cuda(x), want=('prob',), dropout=cfg)['prob'].cpu().numpy()
    assert not np.array_equal(a, c)
