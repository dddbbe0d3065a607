"""Oracle: Dirichlet parameter estimation on the host, float64
(TEST INFRASTRUCTURE ONLY, see oracle/__init__).

Literal loop-level restatement of
  xview/models/dirichletDifferentiation.py:38-57   objective + gradient
  xview/models/dirichletDifferentiation.py:61-99   Hessian pieces, Newton / log-space steps
  xview/models/dirichletDifferentiation.py:129-192 findDirichletPriors
  xview/models/dirichlet_mix.py:207-257            _fit_sufficient_statistic
  xview/models/dirichlet_fastfit.py:188-204,376-395 moment init, fixed point, inverse psi
PINNED against outputs of those reference files executed in the build container
(tests/golden/make_golden.py -> tests/golden/dirichlet_fit.npz).
"""
import math

import numpy as np
from scipy.special import gammaln, polygamma, psi


def _loss(alphas, ss, not_ss, beta, delta):
    """-logProbForMultinomials, dirichletDifferentiation.py:38-45,102-103."""
    for a in alphas:                                     # testTrialPriors :115-120
        if a <= 0:
            return float('inf')
    alphas = np.asarray(alphas, np.float64)
    val = (1 - beta) * gammaln(np.sum(alphas))
    val -= (1 - beta) * np.sum(gammaln(alphas))
    val += np.sum(np.multiply(alphas, ss))
    val -= delta * np.square(alphas).sum()
    val -= beta * np.sum(np.multiply(alphas, not_ss))
    return -1 * val


def _gradient(alphas, ss, not_ss, beta, delta):
    """dirichletDifferentiation.py:48-57."""
    k_dim = len(alphas)
    const = (1 - beta) * psi(sum(alphas))
    grad = [const] * k_dim
    for k in range(k_dim):
        grad[k] += ss[k] - (1 - beta) * psi(alphas[k])
        grad[k] -= 2 * delta * alphas[k]
        grad[k] -= beta * not_ss[k]
    return grad


def _newton_step(alphas, gradient, beta):
    """dirichletDifferentiation.py:61-79 (Minka eq. 18 with constant + diagonal Hessian)."""
    h_const = -(1 - beta) * polygamma(1, sum(alphas))
    h_diag = [(1 - beta) * polygamma(1, a) for a in alphas]
    k_dim = len(gradient)
    num = 0.0
    for i in range(k_dim):
        num += gradient[i] / h_diag[i]
    den = 0.0
    for i in range(k_dim):
        den += 1.0 / h_diag[i]
    b = num / ((1.0 / h_const) + den)
    return [(b - gradient[i]) / h_diag[i] for i in range(k_dim)], h_const, h_diag


def _log_space_step(alphas, gradient, h_const, h_diag):
    """dirichletDifferentiation.py:82-99."""
    k_dim = len(gradient)
    z = 0
    for k in range(k_dim):
        z += alphas[k] / (gradient[k] - alphas[k] * h_diag[k])
    z *= h_const
    ss_ = [1.0 / (gradient[k] - alphas[k] * h_diag[k]) / (1 + z) for k in range(k_dim)]
    s = sum(ss_)
    return [gradient[i] / (gradient[i] - alphas[i] * h_diag[i]) * (1 - h_const * alphas[i] * s)
            for i in range(k_dim)]


def find_dirichlet_priors(ss, not_ss, init_alphas, max_iter=1000, delta=1e-2, beta=1e-2):
    """dirichletDifferentiation.py:129-192, including its control-flow quirks: the
    log-space trial is evaluated (and may abort on overflow) but its result is discarded
    in favour of the gradient back-off loop."""
    priors = init_alphas
    current = _loss(priors, ss, not_ss, beta, delta)
    grad_tol_sq = 2 ** -20
    rate_tol = 2 ** -10
    count = 0
    while count < max_iter:
        count += 1
        gradient = _gradient(priors, ss, not_ss, beta, delta)
        if sum(g ** 2 for g in gradient) < grad_tol_sq:
            return priors
        step, h_const, h_diag = _newton_step(priors, gradient, beta)
        trial = [priors[i] + step[i] for i in range(len(priors))]
        loss = _loss(trial, ss, not_ss, beta, delta)
        if loss < current:
            current = loss
            priors = trial
            continue
        try:
            step = _log_space_step(priors, gradient, h_const, h_diag)
            trial = [priors[i] * math.exp(step[i]) for i in range(len(priors))]
            _loss(trial, ss, not_ss, beta, delta)
        except OverflowError:
            return priors
        loss = 10000000
        rate = 1.0
        while loss > current:
            rate *= 0.9
            trial = [priors[i] + gradient[i] * rate for i in range(len(priors))]
            loss = _loss(trial, ss, not_ss, beta, delta)
        if rate < rate_tol:
            return priors
        current = loss
        priors = trial
    return priors


def fit_sufficient_statistic(counts, class_counts, delta, beta, max_iter=10000):
    """dirichlet_mix.py:207-257 for one modality: counts [C_gt, C_out] (float64 sums of
    log(1e-10+p)), class_counts [C]; returns params [C_out, C_gt] float64."""
    num_classes = len(class_counts)
    params = np.ones((num_classes, num_classes)).astype('float64')
    for c in range(num_classes):
        if class_counts[c] == 0:
            params[:, c] = np.ones(num_classes)
            continue
        ss = (counts[c, :] / class_counts[c]).astype('float64')
        neg_ss = (counts.sum(0) - counts[c, :]) / (class_counts.sum() - class_counts[c])
        prior = np.ones(num_classes).astype('float64')
        params[:, c] = find_dirichlet_priors(ss, neg_ss, prior, max_iter=max_iter,
                                             delta=delta, beta=beta)
    return params


# ---------------------------------------------------------------- Minka fastfit pieces
EULER = -1 * psi(1)


def init_a_moments(d):
    """dirichlet_fastfit.py:376-380: moment-matching initial guess; d is [T, C]."""
    e = d.mean(axis=0)
    e2 = (d ** 2).mean(axis=0)
    return ((e[0] - e2[0]) / (e2[0] - e[0] ** 2)) * e


def ipsi(y, tol=1.48e-9, maxiter=10):
    """dirichlet_fastfit.py:382-395: inverse digamma by Newton iterations."""
    y = np.asanyarray(y, dtype='float')
    x0 = np.where(y >= -2.22, np.exp(y) + 0.5, -1 / (y + EULER))
    for _ in range(maxiter):
        x1 = x0 - (psi(x0) - y) / polygamma(1, x0)
        if np.linalg.norm(x1 - x0) < tol:
            return x1
        x0 = x1
    raise Exception('Unable to converge in {} iterations, value is {}'.format(maxiter, x1))


def fixedpoint_fit(d, tol=1e-7, maxiter=1000):
    """dirichlet_fastfit.py:188-204: Minka fixed point on the rows of d ([T, C])."""
    n = d.shape[0]
    logp = np.log(d).mean(axis=0)

    def ll(a):
        return n * (gammaln(a.sum()) - gammaln(a).sum() + ((a - 1) * logp).sum())

    a0 = init_a_moments(d)
    for _ in range(maxiter):
        a1 = ipsi(psi(a0.sum()) + logp)
        if abs(ll(a1) - ll(a0)) < tol:
            return a1
        a0 = a1
    raise Exception('Failed to converge after {} iterations'.format(maxiter))
