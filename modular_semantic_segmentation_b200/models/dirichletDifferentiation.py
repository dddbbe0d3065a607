"""Host-side float64 fit of a regularised, contrastive Dirichlet likelihood - the B200
package's own statement of xview/models/dirichletDifferentiation.py:129-192
(`findDirichletPriors`), which the reference also runs on the host (dirichlet_mix.py:243-245).

Objective (dirichletDifferentiation.py:38-45), for concentration a (vector over classes):
    (1-beta) [lnGamma(sum a) - sum lnGamma(a)] + a.ss - delta |a|^2 - beta a.not_ss
Newton steps use the Sherman-Morrison form of the (diagonal + constant) Hessian (Minka 2000,
eq. 18); when a Newton step does not lower the loss the iteration falls back to a gradient
step with geometric back-off.  Tolerances: |grad|^2 < 2^-20, learn rate < 2^-10.
"""
import math

import numpy as np
from scipy.special import gammaln, polygamma, psi


def _neg_log_prob(alphas, ss, not_ss, beta, delta):
    alphas = np.asarray(alphas, dtype=np.float64)
    if (alphas <= 0).any():
        return float('inf')
    value = (1 - beta) * gammaln(np.sum(alphas))
    value -= (1 - beta) * np.sum(gammaln(alphas))
    value += np.sum(np.multiply(alphas, ss))
    value -= delta * np.square(alphas).sum()
    value -= beta * np.sum(np.multiply(alphas, not_ss))
    return -value


def _gradient(alphas, ss, not_ss, beta, delta):
    total = 0.0
    for a in alphas:
        total += a                       # left-to-right like the builtin sum of the reference
    common = (1 - beta) * psi(total)
    grad = []
    for k in range(len(alphas)):
        g = common + (ss[k] - (1 - beta) * psi(alphas[k]))
        g -= 2 * delta * alphas[k]
        g -= beta * not_ss[k]
        grad.append(g)
    return grad, total


def _newton_direction(alphas, grad, alpha_sum, beta):
    h_const = -(1 - beta) * polygamma(1, alpha_sum)
    h_diag = [(1 - beta) * polygamma(1, a) for a in alphas]
    num = 0.0
    den = 0.0
    for g, h in zip(grad, h_diag):
        num += g / h
    for h in h_diag:
        den += 1.0 / h
    b = num / ((1.0 / h_const) + den)
    return [(b - g) / h for g, h in zip(grad, h_diag)], h_const, h_diag


def _log_space_overflows(alphas, grad, h_const, h_diag):
    """The reference evaluates a log-space trial (dirichletDifferentiation.py:82-99,166-172)
    whose only observable effect is to stop the fit when exp() overflows."""
    z = 0
    for a, g, h in zip(alphas, grad, h_diag):
        z += a / (g - a * h)
    z *= h_const
    terms = [1.0 / (g - a * h) / (1 + z) for a, g, h in zip(alphas, grad, h_diag)]
    s = sum(terms)
    try:
        for a, g, h in zip(alphas, grad, h_diag):
            math.exp(g / (g - a * h) * (1 - h_const * a * s))
    except OverflowError:
        return True
    return False


def findDirichletPriors(ss, not_ss, initAlphas, max_iter=1000, delta=1e-2, beta=1e-2):
    """Same signature and result as the reference function of that name."""
    priors = initAlphas
    current = _neg_log_prob(priors, ss, not_ss, beta, delta)
    grad_tol_sq = 2 ** -20
    rate_tol = 2 ** -10
    for _ in range(max_iter):
        grad, alpha_sum = _gradient(priors, ss, not_ss, beta, delta)
        size = 0
        for g in grad:
            size += g ** 2
        if size < grad_tol_sq:
            return priors
        step, h_const, h_diag = _newton_direction(priors, grad, alpha_sum, beta)
        trial = [a + d for a, d in zip(priors, step)]
        loss = _neg_log_prob(trial, ss, not_ss, beta, delta)
        if loss < current:
            current, priors = loss, trial
            continue
        if _log_space_overflows(priors, grad, h_const, h_diag):
            return priors
        loss = 10000000
        rate = 1.0
        while loss > current:
            rate *= 0.9
            trial = [a + g * rate for a, g in zip(priors, grad)]
            loss = _neg_log_prob(trial, ss, not_ss, beta, delta)
        if rate < rate_tol:
            return priors
        current, priors = loss, trial
    return priors
