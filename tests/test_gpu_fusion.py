"""GPU parity: per-pixel fusion + score kernels (through the C ABI) against the oracle."""
import numpy as np
import pytest
import torch

import oracle
from util import assert_labels_match, cuda, softmax_probs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    from modular_semantic_segmentation_b200 import device
    device.init()
    return device


@pytest.mark.parametrize('c,npix', [(12, 5000), (10, 256), (13, 777), (2, 33), (24, 1025)])
def test_softmax_argmax(dev, c, npix):
    rng = np.random.default_rng(c * 1000 + npix)
    score = rng.normal(0, 3, size=(npix, c)).astype(np.float32)
    prob, label = dev.softmax_argmax(cuda(score))
    ref = oracle.softmax(score)
    np.testing.assert_allclose(prob.cpu().numpy(), ref, rtol=2e-6, atol=1e-7)
    assert label.dtype == torch.int64
    assert_labels_match(label.cpu().numpy(), ref, 1e-6)
    # ties resolve to the first maximal index (tf.argmax)
    tie = np.zeros((64, c), np.float32)
    tie[:, [1, c - 1]] = 5.0
    _, lt = dev.softmax_argmax(cuda(tie), label_dtype=torch.uint8)
    assert (lt.cpu().numpy() == 1).all()


@pytest.mark.parametrize('c', [12, 13, 3])
def test_label_only_argmax_equals_argmax_of_probabilities_on_near_ties(dev, c):
    """Label-only kernels derive argmax(softmax(score)) from the scores (csrc/argmax.cuh) and
    evaluate the softmax only on near-ties: on scores whose top classes differ by 0, 1, 2, 4 or 8
    ulps (in either index order) the labels must equal the argmax of the probabilities the
    probability-writing kernel returns, and follow tf.argmax's first-index rule on exact ties."""
    rng = np.random.default_rng(500 + c)
    npix = 40000
    score = rng.normal(0, 2, size=(npix, c)).astype(np.float32)
    top = score.max(-1)
    first = rng.integers(0, c, size=npix)
    second = (first + rng.integers(1, c, size=npix)) % c
    ulps = rng.choice([0, 1, 2, 4, 8], size=npix)
    rows = np.arange(npix)
    score[rows, first] = top
    near = top.copy()
    for _ in range(8):
        step = ulps > 0
        near[step] = np.nextafter(near[step], np.float32(-np.inf))
        ulps = np.maximum(ulps - 1, 0)
    score[rows, second] = near
    d = cuda(score)
    prob, label_full = dev.softmax_argmax(d)
    _, label_only = dev.softmax_argmax(d, want_prob=False, label_dtype=torch.uint8)
    _, label_only64 = dev.softmax_argmax(d, want_prob=False)
    want = prob.cpu().numpy().argmax(-1)
    np.testing.assert_array_equal(label_full.cpu().numpy(), want)
    np.testing.assert_array_equal(label_only.cpu().numpy().astype(np.int64), want)
    np.testing.assert_array_equal(label_only64.cpu().numpy(), want)
    # the construction exercises both outcomes: the earlier near-tied class wins sometimes
    earlier = (second < first)
    assert (want[earlier] == second[earlier]).any() and (want[earlier] == first[earlier]).any()


@pytest.mark.parametrize('prior', ['data', 'uniform', 0.3])
@pytest.mark.parametrize('label_dtype', [torch.int64, torch.uint8])
def test_bayes_fusion_bit_exact(dev, exp868, prior, label_dtype):
    """Integer work: the device lookups must reproduce the oracle argmax exactly, both through
    the decision table and through the literal log-likelihood sum."""
    c = 12
    rng = np.random.default_rng(5)
    cms = [exp868['cm_measure_rgb'].astype('float32').T,
           exp868['cm_measure_depth'].astype('float32').T]
    labels = [rng.integers(0, c, size=(2, 37, 53)) for _ in range(2)]
    score_ref, lls, _ = oracle.bayes_fusion(labels, cms, prior)
    ref = oracle.argmax_first(score_ref)
    # (a) decision table built with the oracle's float32 arithmetic, integer lookups on device
    a, b = np.meshgrid(np.arange(c), np.arange(c), indexing='ij')
    lut = oracle.argmax_first(oracle.bayes_fusion([a, b], cms, prior)[0]).astype(np.int32)
    dl = [cuda(l, label_dtype) for l in labels]
    out = dev.bayes_fuse_lut(dl, cuda(lut), c)
    assert out.dtype == label_dtype
    np.testing.assert_array_equal(out.cpu().numpy().astype(np.int64), ref)
    # (b) literal form: device adds the same float32 table rows in the same order
    with np.errstate(divide='ignore'):
        log_cond = np.stack([np.log(np.float32(1e-20) + oracle.bayes_conditionals(m)) for m in cms])
        log_prior = np.log(np.asarray(oracle.bayes_prior(cms[-1], prior), np.float32))
    log_prior = np.broadcast_to(log_prior, (c,)).astype(np.float32)
    score, out2 = dev.bayes_fuse_score(dl, cuda(log_cond), cuda(log_prior))
    np.testing.assert_array_equal(score.cpu().numpy(), score_ref.astype(np.float32))
    np.testing.assert_array_equal(out2.cpu().numpy().astype(np.int64), ref)


def _dirichlet_tables(params, sigma, prior):
    """alpha - 1, lbeta(alpha) and log(1e-20 + prior) in float32, built like the product's
    dirichlet_tables and oracle.dirichlet_fusion_f32 build them: the normaliser is evaluated in
    float64 on the float32 concentrations and rounded once."""
    alpha = [(np.float32(sigma) * p.astype('float32')) for p in params]
    am1 = np.stack([a - np.float32(1) for a in alpha]).astype(np.float32)
    lognorm = np.stack([oracle.dirichlet_log_norm(a.astype(np.float64)).astype(np.float32)
                        for a in alpha])
    logprior = np.log(np.float32(1e-20) + np.asarray(prior, np.float32)).astype(np.float32)
    return am1, lognorm, logprior


@pytest.mark.parametrize('c', [12, 14, 5])
def test_dirichlet_fusion(dev, c):
    rng = np.random.default_rng(c)
    shape = (2, 40, 31, c)
    probs = [softmax_probs(rng, shape), softmax_probs(rng, shape)]
    params = [1 + rng.gamma(2, 2, size=(c, c)) for _ in range(2)]
    counts = rng.integers(1, 1000, size=c)
    prior = oracle.dirichlet_prior(counts)
    ref64 = oracle.dirichlet_fusion(probs, params, prior, sigma=1.0, dtype=np.float64)
    am1, lognorm, logprior = _dirichlet_tables(params, 1.0, prior)
    score, label = dev.dirichlet_fuse([cuda(p) for p in probs], cuda(am1), cuda(lognorm),
                                      cuda(logprior), want_score=True)
    got = score.cpu().numpy()
    scale = np.abs(ref64).max()
    np.testing.assert_allclose(got, ref64, rtol=0, atol=2e-5 * scale)
    assert_labels_match(label.cpu().numpy(), ref64, 4e-5 * scale)


def test_dirichlet_fusion_exact_on_decisive_inputs(dev):
    """Identical probabilities, decisive margins -> labels identical to the float32 oracle."""
    c = 12
    rng = np.random.default_rng(3)
    shape = (1, 64, 64, c)
    probs = [softmax_probs(rng, shape, 3.0), softmax_probs(rng, shape, 3.0)]
    params = [1 + 8 * np.eye(c) + rng.random((c, c)) for _ in range(2)]
    prior = oracle.dirichlet_prior(np.ones(c))
    ref32 = oracle.dirichlet_fusion(probs, params, prior, dtype=np.float32)
    am1, lognorm, logprior = _dirichlet_tables(params, 1.0, prior)
    _, label = dev.dirichlet_fuse([cuda(p) for p in probs], cuda(am1), cuda(lognorm),
                                  cuda(logprior))
    from util import top2_margin
    decisive = top2_margin(ref32.astype(np.float64)) > 1e-3
    assert decisive.mean() > 0.99
    np.testing.assert_array_equal(label.cpu().numpy()[decisive], np.argmax(ref32, -1)[decisive])


def _flip_report(name, fast, exact, redone, total):
    flips = int((fast != exact).sum())
    print('%s: fast-mode flips %d of %d pixels (%.3g), exact re-evaluations %d (%.3g)'
          % (name, flips, total, flips / total, redone, redone / total))
    return flips


def test_dirichlet_fusion_exact_mode_is_bit_exact_at_full_size(dev):
    """north_star: fusion argmax bit-exact.  16 x 768 x 384 random softmax outputs (the batch of
    BASELINE configs[1]/[2]) through the exact mode: labels AND scores must equal the fixed-order
    float32 oracle (dirichlet_mix.py:14-36,100-113 restated in oracle.dirichlet_fusion_f32) to
    the last bit; the flip rate of the fast arithmetic is printed."""
    c = 12
    rng = np.random.default_rng(2024)
    shape = (16, 768, 384, c)
    probs = [softmax_probs(rng, shape), softmax_probs(rng, shape)]
    params = [1 + rng.gamma(2, 2, size=(c, c)) + 4 * np.eye(c) for _ in range(2)]
    prior = oracle.dirichlet_prior(rng.integers(100, 100000, size=c))
    ref = oracle.dirichlet_fusion_f32(probs, params, prior)
    ref_label = oracle.argmax_first(ref)
    tables = [cuda(t) for t in _dirichlet_tables(params, 1.0, prior)]
    dp = [cuda(p) for p in probs]
    redone = torch.zeros(1, dtype=torch.int64, device='cuda')
    score, label_all = dev.dirichlet_fuse(dp, *tables, want_score=True, exact=True)
    np.testing.assert_array_equal(score.cpu().numpy(), ref)
    np.testing.assert_array_equal(label_all.cpu().numpy(), ref_label)
    _, label = dev.dirichlet_fuse(dp, *tables, exact=True, num_exact=redone)
    np.testing.assert_array_equal(label.cpu().numpy(), ref_label)
    _, fast = dev.dirichlet_fuse(dp, *tables, label_dtype=torch.uint8)
    npix = ref_label.size
    _flip_report('dirichlet C=12 16x768x384', fast.cpu().numpy(), ref_label, int(redone.item()), npix)
    # the two-tier scheme must re-evaluate only a small share of the pixels
    assert int(redone.item()) < 0.05 * npix


@pytest.mark.parametrize('c', [12, 13, 5])
def test_dirichlet_fusion_exact_mode_on_adversarial_near_ties(dev, c):
    """Class columns that are exact copies of each other (exact ties: the first index must
    win), copies nudged by one ulp, and pixels that are duplicates of each other up to one ulp
    of one probability: the exact mode must still reproduce the oracle argmax bit for bit."""
    rng = np.random.default_rng(70 + c)
    shape = (3, 96, 80, c)
    probs = [softmax_probs(rng, shape, 1.0), softmax_probs(rng, shape, 1.0)]
    # near-duplicate pixels: pixel 2i+1 = pixel 2i with one probability moved by one ulp
    for p in probs:
        flat = p.reshape(-1, c)
        flat[1::2] = flat[0::2]
        idx = rng.integers(0, c, size=flat[1::2].shape[0])
        rows = np.arange(1, flat.shape[0], 2)
        flat[rows, idx] = np.nextafter(flat[rows, idx], np.float32(1))
    params = [1 + rng.gamma(2, 2, size=(c, c)) for _ in range(2)]
    for a in params:
        a[:, 2] = a[:, 0]                                   # exact tie between classes 0 and 2
        a[:, 3] = a[:, 1]
        a[0, 3] = np.nextafter(np.float32(a[0, 3]), np.float32(100))   # 1-ulp near tie 1 vs 3
        if c > 4:
            a[:, 4] = a[:, 0] * (1 + 1e-6)
    counts = np.full(c, 50.0)
    prior = oracle.dirichlet_prior(counts)                  # equal priors keep the ties exact
    ref = oracle.dirichlet_fusion_f32(probs, params, prior)
    ref_label = oracle.argmax_first(ref)
    tables = [cuda(t) for t in _dirichlet_tables(params, 1.0, prior)]
    dp = [cuda(p) for p in probs]
    redone = torch.zeros(1, dtype=torch.int64, device='cuda')
    score, _ = dev.dirichlet_fuse(dp, *tables, want_score=True, exact=True)
    np.testing.assert_array_equal(score.cpu().numpy(), ref)
    for dtype in (torch.int64, torch.uint8):
        _, label = dev.dirichlet_fuse(dp, *tables, exact=True, label_dtype=dtype, num_exact=redone)
        np.testing.assert_array_equal(label.cpu().numpy().astype(np.int64), ref_label)
    # the construction really produces ties: classes 0 and 2 have identical scores everywhere
    np.testing.assert_array_equal(ref[..., 0], ref[..., 2])
    assert not (ref_label == 2).any()
    _, fast = dev.dirichlet_fuse(dp, *tables)
    _flip_report('dirichlet adversarial C=%d' % c, fast.cpu().numpy(), ref_label,
                 int(redone.item()) // 2, ref_label.size)


def test_average_and_variance_fusion_bit_exact_at_full_size(dev):
    """average_mix.py:18-21 and variance_mix.py:7-15 on 16 x 768 x 384 probabilities: every device
    operation is individually rounded in numpy's order, so scores and labels are bit-exact."""
    c = 12
    rng = np.random.default_rng(99)
    shape = (16, 768, 384, c)
    probs = [softmax_probs(rng, shape), softmax_probs(rng, shape)]
    probs[1][0, :8] = probs[0][0, :8]                       # rows of exact ties after averaging
    dp = [cuda(p) for p in probs]
    ref = oracle.average_fusion(probs)
    assert ref.dtype == np.float32
    score, label = dev.average_fuse(dp, want_score=True)
    np.testing.assert_array_equal(score.cpu().numpy(), ref)
    np.testing.assert_array_equal(label.cpu().numpy(), oracle.argmax_first(ref))
    variances = [(rng.random(shape[:-1] + (1,)) * 1e-2).astype(np.float32) for _ in range(2)]
    variances[0][0, 0, :5] = 0.0
    refv = oracle.variance_fusion(probs, variances)
    assert refv.dtype == np.float32
    score, label = dev.variance_fuse(dp, [cuda(v[..., 0]) for v in variances], want_score=True)
    np.testing.assert_array_equal(score.cpu().numpy(), refv)
    np.testing.assert_array_equal(label.cpu().numpy(), oracle.argmax_first(refv))


def test_average_and_variance_fusion(dev):
    c = 12
    rng = np.random.default_rng(11)
    shape = (2, 33, 47, c)
    probs = [softmax_probs(rng, shape), softmax_probs(rng, shape)]
    ref = oracle.average_fusion(probs)
    score, label = dev.average_fuse([cuda(p) for p in probs], want_score=True)
    np.testing.assert_allclose(score.cpu().numpy(), ref, rtol=1e-6, atol=1e-8)
    assert_labels_match(label.cpu().numpy(), ref, 1e-6)
    variances = [rng.random(shape[:-1] + (1,)).astype(np.float32) * 1e-2 for _ in range(2)]
    variances[0][0, 0, :5] = 0.0           # exercises the 1e-20 guard
    refv = oracle.variance_fusion(probs, variances)
    score, label = dev.variance_fuse([cuda(p) for p in probs],
                                     [cuda(v[..., 0]) for v in variances], want_score=True)
    np.testing.assert_allclose(score.cpu().numpy(), refv, rtol=2e-6, atol=1e-8)
    assert_labels_match(label.cpu().numpy(), refv, 1e-6)


def test_mc_moments(dev):
    c, t = 12, 20
    rng = np.random.default_rng(7)
    samples = softmax_probs(rng, (t, 2, 21, 35, c))
    mean_ref, unc = oracle.sampling_uncertainty(samples.astype(np.float64))
    _, var_ref = oracle.mc_moments(samples.astype(np.float64), 0)
    out = dev.mc_moments(cuda(samples), want=('mean', 'var', 'mean_var', 'entropy',
                                               'cond_entropy', 'sum_var'))
    np.testing.assert_allclose(out['mean'].cpu().numpy(), mean_ref, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out['var'].cpu().numpy(), var_ref, rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(out['mean_var'].cpu().numpy(), var_ref.mean(-1), rtol=1e-4,
                               atol=1e-8)
    np.testing.assert_allclose(out['sum_var'].cpu().numpy(), unc['variance'], rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(out['entropy'].cpu().numpy(), unc['entropy'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out['cond_entropy'].cpu().numpy(), unc['cond_entropy'], rtol=1e-4,
                               atol=1e-6)


def test_sufficient_statistics(dev):
    c = 12
    rng = np.random.default_rng(13)
    prob = softmax_probs(rng, (3, 40, 56, c))
    labels = rng.integers(-1, c, size=(3, 40, 56)).astype(np.int32)
    s_ref, n_ref = oracle.sufficient_statistics(prob, labels, c)
    stats = torch.zeros((c, c), dtype=torch.float64, device='cuda')
    counts = torch.zeros(c, dtype=torch.int64, device='cuda')
    dev.dirichlet_suffstats(cuda(prob), cuda(labels), stats, counts)
    dev.dirichlet_suffstats(cuda(prob), cuda(labels), stats, counts)   # accumulates
    np.testing.assert_array_equal(counts.cpu().numpy(), 2 * n_ref)
    np.testing.assert_allclose(stats.cpu().numpy(), 2 * s_ref, rtol=2e-6)


@pytest.mark.parametrize('pred_dtype', [torch.int64, torch.uint8])
def test_confusion_matrix_bit_exact(dev, pred_dtype):
    c = 12
    rng = np.random.default_rng(17)
    labels = rng.integers(-1, c, size=(4, 96, 128)).astype(np.int32)
    labels[0, :10] = 3                      # long same-class runs (warp aggregation path)
    pred = rng.integers(0, c, size=labels.shape)
    pred[0, :10] = 3
    ref = oracle.confusion_matrix(labels, pred, c)
    cm = torch.zeros((c, c), dtype=torch.int64, device='cuda')
    dev.confusion_accumulate(cuda(pred, pred_dtype), cuda(labels), cm)
    np.testing.assert_array_equal(cm.cpu().numpy(), ref)
    dev.confusion_accumulate(cuda(pred, pred_dtype), cuda(labels), cm)
    np.testing.assert_array_equal(cm.cpu().numpy(), 2 * ref)
    assert cm.sum().item() == 2 * int((labels >= 0).sum())


def test_empty_and_ragged_inputs(dev):
    c = 12
    cm = torch.zeros((c, c), dtype=torch.int64, device='cuda')
    empty_pred = torch.zeros((0,), dtype=torch.int64, device='cuda')
    empty_lab = torch.zeros((0,), dtype=torch.int32, device='cuda')
    dev.confusion_accumulate(empty_pred, empty_lab, cm)
    assert cm.sum().item() == 0
    rng = np.random.default_rng(1)
    for npix in (1, 255, 257):              # around the 256-pixel tile size
        score = rng.normal(size=(npix, c)).astype(np.float32)
        prob, label = dev.softmax_argmax(cuda(score))
        np.testing.assert_allclose(prob.cpu().numpy(), oracle.softmax(score), rtol=2e-6, atol=1e-7)


def test_per_pixel_dirichlet_fit_over_mc_samples(dev):
    """a15: batched moment-init + fixed-point Dirichlet MLE; every pixel against the float64
    oracle (dirichlet_fastfit restatement, itself pinned to the reference's outputs)."""
    c, t = 5, 20
    rng = np.random.default_rng(23)
    true_alpha = rng.gamma(2.0, 2.0, size=(6, 7, c)) + 0.5
    samples = np.stack([np.stack([rng.dirichlet(true_alpha[i, j], size=t) for j in range(7)])
                        for i in range(6)])                        # [6,7,T,C]
    samples = np.ascontiguousarray(np.moveaxis(samples, 2, 0)).astype(np.float32)   # [T,6,7,C]
    alpha, iters = dev.dirichlet_fit_samples(cuda(samples), tol=1e-6, maxiter=200,
                                             want_iterations=True)
    alpha = alpha.cpu().numpy()
    assert alpha.shape == (6, 7, c) and (iters.cpu().numpy() >= 1).all()
    for i in range(6):
        for j in range(7):
            ref = oracle.fixedpoint_fit(samples[:, i, j].astype(np.float64), tol=1e-9,
                                        maxiter=5000)
            np.testing.assert_allclose(alpha[i, j], ref, rtol=2e-2)
    # the moment initialisation alone (maxiter=0) equals dirichlet_fastfit._init_a
    a0 = dev.dirichlet_fit_samples(cuda(samples), maxiter=0).cpu().numpy()
    np.testing.assert_allclose(a0[2, 3], oracle.init_a_moments(samples[:, 2, 3].astype(np.float64)),
                               rtol=2e-3)


def test_dirichlet_uncertainty_fusion(dev):
    c = 6
    rng = np.random.default_rng(29)
    shape = (2, 17, 23, c)
    probs = [softmax_probs(rng, shape), softmax_probs(rng, shape)]
    variances = [(rng.random(shape) * 1e-2).astype(np.float32) for _ in range(2)]
    cond = [1 + rng.gamma(2, 2, size=(c, c)).astype(np.float32) for _ in range(2)]
    prior = oracle.dirichlet_prior(rng.integers(1, 100, size=c))
    ref = oracle.dirichlet_uncertainty_fusion(probs, cond, variances, prior)
    logprior = np.log(np.float32(1e-20) + prior).astype(np.float32)
    score, label = dev.dirichlet_uncertainty_fuse([cuda(p) for p in probs],
                                                  [cuda(v) for v in variances],
                                                  cuda(np.stack(cond)), cuda(logprior),
                                                  want_score=True)
    scale = np.abs(ref).max()
    np.testing.assert_allclose(score.cpu().numpy(), ref, rtol=0, atol=1e-4 * scale)
    assert_labels_match(label.cpu().numpy(), ref, 2e-4 * scale)
    # the reference-named entry point (uncertainty_dirichlet_mix.py:18) over the same kernel
    from xview.models.uncertainty_dirichlet_mix import dirichlet_uncertainty_fusion
    again = dirichlet_uncertainty_fusion([cuda(p) for p in probs], cond,
                                         [cuda(v) for v in variances], prior)
    assert torch.equal(again, score)
