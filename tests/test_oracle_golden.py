"""Pin the oracle against every known-answer vector the reference holds for the path
(SURVEY.md section 8c).  CPU only."""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN


@pytest.mark.parametrize('key', ['fusion', 'rgb', 'depth'])
def test_score_measures_match_stored_run_868(exp868, key):
    """base_model.py:315-329 recomputed from the stored matrices reproduces the stored
    measures (scalars to the last bit, arrays to their printed 8 digits)."""
    m = oracle.score_measures(exp868['cm_test_%s' % key])
    for scalar in ('mean_IoU', 'total_accuracy', 'mean_F1'):
        assert m[scalar] == float(exp868['%s_%s' % (key, scalar)]), scalar
    for arr in ('IoU', 'F1', 'precision', 'recall'):
        np.testing.assert_allclose(m[arr], exp868['%s_%s' % (key, arr)], rtol=0, atol=5e-9)


def test_published_numbers(exp868):
    assert float(exp868['fusion_mean_IoU']) == 0.6877204624612501
    assert float(exp868['fusion_total_accuracy']) == 0.920302592013555
    # 149 frames of 768x384 in the measure set
    assert exp868['cm_measure_rgb'].sum() == exp868['cm_measure_depth'].sum() == 43941888


def test_bayes_fusion_consistent_with_decision_matrix(exp868):
    """The two reference statements of the rule (bayes_mix.py:12-58 per pixel, :61-112 as a
    C^M lookup table) agree on the real matrices of run 868 for every label pair."""
    cms64 = [exp868['cm_measure_rgb'].T, exp868['cm_measure_depth'].T]
    c = 12
    a, b = np.meshgrid(np.arange(c), np.arange(c), indexing='ij')
    for prior in ('data', 'uniform', 0.3):
        lut = oracle.bayes_decision_matrix(cms64, prior)
        score, lls, conds = oracle.bayes_fusion([a[None], b[None]], cms64, prior)
        assert score.shape == (1, c, c, c)
        np.testing.assert_array_equal(oracle.argmax_first(score)[0], lut)
        assert len(lls) == 2 and conds[0].shape == (1, c, c, c)
    # float32 tables as BayesFusion builds them (bayes_mix.py:141)
    cms32 = [m.astype('float32') for m in cms64]
    score32, _, _ = oracle.bayes_fusion([a[None], b[None]], cms32, 'data')
    assert score32.dtype == np.float32
    assert (oracle.argmax_first(score32)[0] == oracle.bayes_decision_matrix(cms64)).mean() > 0.97


def test_bayes_zero_count_class_is_never_selected():
    rng = np.random.default_rng(0)
    cm = rng.integers(1, 50, size=(6, 6)).astype('float32')
    cm[:, 2] = 0            # gt class 2 never observed -> conditional nan->0, prior 0
    lut = oracle.bayes_decision_matrix([cm, cm])
    assert not (lut == 2).any()


def test_bilinear_kernel_closed_form():
    np.testing.assert_allclose(oracle.bilinear_kernel_1d(4), [.25, .75, .75, .25])
    k16 = oracle.bilinear_kernel_1d(16)
    np.testing.assert_allclose(k16[:8], (2 * np.arange(8) + 1) / 16.0)
    np.testing.assert_allclose(k16, k16[::-1])
    w = oracle.bilinear_filter((4, 4, 3, 3))
    assert w.shape == (4, 4, 3, 3) and w[1, 2, 0, 0] == np.float32(.75 * .75)
    assert w[:, :, 0, 1].sum() == 0


def test_weight_key_list_matches_reference_printout():
    keys = json.load(open(os.path.join(GOLDEN, 'fcn_weight_keys.json')))
    for prefix, cin in (('rgb', 3), ('depth', 1)):
        shapes = oracle.fcn_param_shapes(prefix, cin, 64, 12)
        assert list(shapes) == keys[prefix]
        assert len(shapes) == 34
    assert oracle.fcn_param_shapes('rgb', 3, 64, 12)['rgb/upscore/kernel'] == (16, 16, 64, 64)


def test_dirichlet_fit_matches_reference_outputs(dirichlet_golden):
    g = dirichlet_golden
    for i in g['fit_cases']:
        delta, beta = g['fit%d_delta_beta' % i]
        c = len(g['fit%d_ss' % i])
        alpha = oracle.find_dirichlet_priors(g['fit%d_ss' % i], g['fit%d_neg_ss' % i],
                                             np.ones(c), max_iter=10000, delta=delta,
                                             beta=beta)
        np.testing.assert_allclose(np.asarray(alpha, np.float64), g['fit%d_alpha' % i],
                                   rtol=1e-12, atol=0)


def test_fastfit_pieces_match_reference_outputs(dirichlet_golden):
    g = dirichlet_golden
    np.testing.assert_allclose(oracle.init_a_moments(g['ff_D']), g['ff_init_a'], rtol=1e-13)
    np.testing.assert_allclose(oracle.ipsi(g['ff_ipsi_y']), g['ff_ipsi_x'], rtol=1e-13)
    np.testing.assert_allclose(oracle.fixedpoint_fit(g['ff_D']), g['ff_fixedpoint'],
                               rtol=1e-12)


def test_confusion_matrix_ignores_negative_labels():
    labels = np.array([[0, 1, -1, 2], [2, 2, -5, 1]])
    pred = np.array([[0, 2, 1, 2], [2, 0, 0, 1]])
    cm = oracle.confusion_matrix(labels, pred, 3)
    assert cm.dtype == np.int64 and cm.sum() == 6
    np.testing.assert_array_equal(cm, [[1, 0, 0], [0, 1, 1], [1, 0, 2]])


def test_fcn_shapes_and_relu_identity_of_bilinear_upscore():
    rng = np.random.default_rng(1)
    params = oracle.glorot_fcn_params('rgb', 3, 8, 5, rng)
    x = rng.uniform(0, 1, size=(1, 32, 48, 3)).astype(np.float32)
    out = oracle.test_pipeline(x, params, 'rgb', 8, 5)
    assert out['fused'].shape == (1, 4, 6, 8)
    assert out['score'].shape == (1, 32, 48, 5)
    assert out['classification'].dtype == np.int64
    np.testing.assert_allclose(out['prob'].sum(-1), 1, rtol=1e-5)
    # the decoder is linear in `fused` for the bilinear kernel: score 1x1 and x8 upscore commute
    lowres = oracle.conv2d(out['fused'], params, 'rgb/score', activation=False)
    p2 = dict(params)
    p2['rgb/upscore/kernel'] = oracle.bilinear_filter((16, 16, 5, 5))
    up = oracle.deconv2d(lowres - params['rgb/score/bias'], p2, 'rgb/upscore', 8,
                         activation=False) + params['rgb/score/bias']
    np.testing.assert_allclose(up, out['score'], rtol=1e-4, atol=1e-5)


def test_adapnet_oracle_geometry():
    """TF 'SAME' padding rule (smaller half first) and the layer shapes of adapnet.py:99-173."""
    from oracle.adapnet import same_padding
    assert same_padding(6, 7, 2, 1) == (2, 3)      # 7x7 stride 2 on an even size
    assert same_padding(5, 3, 2, 1) == (1, 1)
    assert same_padding(8, 1, 2, 1) == (0, 0)      # 1x1 stride 2 picks pixels 0, 2, 4, ...
    assert same_padding(8, 3, 1, 4) == (4, 4)      # atrous 3x3 rate 4
    assert same_padding(7, 2, 2, 1) == (0, 1)
    # strided case by hand: x = 0..5, w = ones(3), stride 2 -> out = 3, pad_total = 1 -> (0, 1):
    # windows [0,1,2], [2,3,4], [4,5,pad] -> 3, 9, 9
    assert same_padding(6, 3, 2, 1) == (0, 1)
    x = np.arange(6, dtype=np.float32).reshape(1, 6, 1, 1)
    kernel = np.zeros((3, 3, 1, 1), np.float32)
    kernel[:, 1] = 1                               # acts along the height axis only
    unit = {'s/kernel': kernel, 's/gamma': np.ones(1, np.float32),
            's/beta': np.zeros(1, np.float32), 's/moving_mean': np.zeros(1, np.float32),
            's/moving_variance': np.full(1, 1 - 1e-3, np.float32)}
    y = oracle.conv_bn(x, unit, 's', stride=2)
    np.testing.assert_allclose(y[0, :, 0, 0], [3.0, 9.0, 9.0], rtol=1e-6)
    rng = np.random.default_rng(0)
    p = oracle.adapnet_params('rgb', 3, 20, 14, rng)
    assert len(p) == 334 and p['rgb/first_deconvolution_upconv/kernel'].shape == (4, 4, 20, 2048)
    assert 'rgb/block_layer_1/stage_1/bias' not in p and 'rgb/shortcut/bias' in p
    out = oracle.adapnet(rng.normal(size=(1, 32, 48, 3)).astype(np.float32), p, 'rgb', 20, 14)
    assert out['block_0_2'].shape == (1, 16, 24, 64) and out['block_3'].shape == (1, 8, 12, 256)
    assert out['block_7'].shape == (1, 4, 6, 512) and out['block_16'].shape == (1, 2, 3, 2048)
    assert out['merge'].shape == (1, 4, 6, 20) and out['score'].shape == (1, 32, 48, 14)


def test_oracle_layer_ops_match_their_definitions():
    """The torch-CPU calls inside the oracle against direct numpy loops over the documented
    tf.layers definitions: 'same' conv as a correlation with symmetric
    zero padding, conv2d_transpose 'SAME' as out[s*i + k - (K-s)/2] += x[i] * w[k] with the
    [kh,kw,Cout,Cin] kernel and no flip, 2x2/2 'valid' max pooling, softmax over the last axis
    and first-index argmax."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((1, 5, 6, 3))
    w = rng.standard_normal((3, 3, 3, 4))
    b = rng.standard_normal(4)
    got = oracle.conv2d(x, {'l/kernel': w, 'l/bias': b}, 'l', activation=False)
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    want = np.zeros((1, 5, 6, 4))
    for y in range(5):
        for xx in range(6):
            want[0, y, xx] = np.einsum('abi,abio->o', xp[0, y:y + 3, xx:xx + 3], w) + b
    np.testing.assert_allclose(got, want, atol=1e-12)

    for k, s in ((4, 2), (16, 8)):
        xin = rng.standard_normal((1, 3, 2, 2))
        wt = rng.standard_normal((k, k, 3, 2))           # [kh, kw, Cout, Cin]
        got = oracle.deconv2d(xin, {'l/kernel': wt}, 'l', s, activation=False)
        pad = (k - s) // 2
        want = np.zeros((1, 3 * s, 2 * s, 3))
        for iy in range(3):
            for ix in range(2):
                for ky in range(k):
                    for kx in range(k):
                        oy, ox = s * iy + ky - pad, s * ix + kx - pad
                        if 0 <= oy < 3 * s and 0 <= ox < 2 * s:
                            want[0, oy, ox] += wt[ky, kx] @ xin[0, iy, ix]
        np.testing.assert_allclose(got, want, atol=1e-12)

    p = oracle.max_pool2x2(x[:, :4])
    assert p.shape == (1, 2, 3, 3) and p[0, 1, 2, 1] == x[0, 2:4, 4:6, 1].max()
    sm = oracle.softmax(x)
    np.testing.assert_allclose(sm, np.exp(x) / np.exp(x).sum(-1, keepdims=True), atol=1e-12)
    tie = np.array([[1.0, 3.0, 3.0, 2.0]])
    assert oracle.argmax_first(tie)[0] == 1

    # Adapnet's strided / atrous convolution with TF 'SAME' (smaller padding half first)
    xs = rng.standard_normal((1, 6, 8, 2))
    ws = rng.standard_normal((3, 3, 2, 3))
    unit = {'s/kernel': ws, 's/gamma': np.ones(3), 's/beta': np.zeros(3),
            's/moving_mean': np.zeros(3), 's/moving_variance': np.full(3, 1 - 1e-3)}
    for stride, dil in ((2, 1), (1, 2)):
        got = oracle.conv_bn(xs, unit, 's', stride=stride, dilation=dil, activation=False)
        ho, wo = -(-6 // stride), -(-8 // stride)
        pt = max((ho - 1) * stride + 2 * dil + 1 - 6, 0) // 2
        pl = max((wo - 1) * stride + 2 * dil + 1 - 8, 0) // 2
        want = np.zeros((1, ho, wo, 3))
        for oy in range(ho):
            for ox in range(wo):
                for ky in range(3):
                    for kx in range(3):
                        yy, xx = oy * stride + ky * dil - pt, ox * stride + kx * dil - pl
                        if 0 <= yy < 6 and 0 <= xx < 8:
                            want[0, oy, ox] += xs[0, yy, xx] @ ws[ky, kx]
        np.testing.assert_allclose(got, want, atol=1e-10)
