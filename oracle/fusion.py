"""Oracle: per-pixel fusion rules (TEST INFRASTRUCTURE ONLY, see oracle/__init__).

numpy restatement of
  xview/models/bayes_mix.py:12-58      bayes_fusion
  xview/models/bayes_mix.py:61-112     bayes_decision_matrix
  xview/models/dirichlet_mix.py:14-36  dirichlet_fusion (+ graph side :96-136)
  xview/models/dirichlet_mix.py:142-163 sufficient statistics
  xview/models/average_mix.py:18-21    average fusion
  xview/models/variance_mix.py:7-15    variance_fusion
  xview/models/variance_mix.py:62-66, bayesian_fcn.py:48-57, custom_layers.py:251-256
                                       MC-dropout moments / uncertainties
Arithmetic dtype follows the arrays handed in (float32 through the model classes,
float64 when raw record arrays are passed as experiments/timing.py:61-68 does).
"""
from itertools import product

import numpy as np
from scipy.special import gammaln


# ----------------------------------------------------------------------------- Bayes
def bayes_conditionals(confusion_matrix):
    """bayes_mix.py:35: p(expert output | gt class) = column-normalised matrix, nan -> 0.
    `confusion_matrix` is what bayes_fusion receives: rows = expert output, cols = gt class
    (BayesFusion transposes score()'s matrix at bayes_mix.py:141)."""
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.nan_to_num(confusion_matrix / confusion_matrix.sum(0))


def bayes_prior(confusion_matrix_last, class_prior='data'):
    """bayes_mix.py:42-54.  The uniform prior is the constant 1/14 regardless of C and the
    data prior comes from the LAST expert's matrix (reference quirks)."""
    uniform_prior = 1.0 / 14
    with np.errstate(divide='ignore', invalid='ignore'):
        data_prior = confusion_matrix_last.sum(0) / confusion_matrix_last.sum()
    if class_prior == 'uniform':
        return uniform_prior
    if class_prior == 'data':
        return data_prior
    weight = float(class_prior)
    prior = weight * uniform_prior + (1 - weight) * data_prior
    return prior / prior.sum()


def bayes_fusion(classifications, confusion_matrices, class_prior='data'):
    """bayes_mix.py:12-58.  classifications: list of int arrays [N,H,W]; returns
    (score [N,H,W,C], log_likelihoods list, conditionals list)."""
    dtype = np.asarray(confusion_matrices[0]).dtype
    log_likelihoods, conditionals = [], []
    for i_expert in range(len(confusion_matrices)):
        conditional = bayes_conditionals(np.asarray(confusion_matrices[i_expert]))
        gathered = conditional[classifications[i_expert]]  # tf.gather on axis 0
        conditionals.append(gathered)
        with np.errstate(divide='ignore'):
            log_likelihoods.append(np.log(np.asarray(1e-20, dtype) + gathered))
    prior = bayes_prior(np.asarray(confusion_matrices[-1]), class_prior)
    with np.errstate(divide='ignore'):
        log_prior = np.log(np.asarray(prior, dtype))
    total = log_likelihoods[0]
    for ll in log_likelihoods[1:]:
        total = total + ll
    return total + log_prior, log_likelihoods, conditionals


def bayes_decision_matrix(confusion_matrices, class_prior='data'):
    """bayes_mix.py:61-112: the same rule for all C^M label combinations; note
    log_likelihoods is allocated float64 (:85) whatever dtype the matrices have."""
    num_classes = confusion_matrices[0].shape[0]
    num_experts = len(confusion_matrices)
    combos = np.array(list(product(*(range(num_classes) for _ in range(num_experts)))))
    log_likelihoods = np.zeros((combos.shape[0], num_experts, num_classes))
    for i_expert in range(num_experts):
        conditional = bayes_conditionals(np.asarray(confusion_matrices[i_expert]))
        with np.errstate(divide='ignore'):
            log_likelihoods[:, i_expert, :] = np.log(1e-20 + conditional[combos[:, i_expert]])
    prior = bayes_prior(np.asarray(confusion_matrices[-1]), class_prior)
    with np.errstate(divide='ignore'):
        fused = np.argmax(log_likelihoods.sum(1) + np.log(prior), axis=1)
    return fused.reshape([num_classes for _ in range(num_experts)])


# ------------------------------------------------------------------------- Dirichlet
def dirichlet_log_norm(alpha):
    """lbeta(alpha) over axis 0 = sum lgamma(alpha_k) - lgamma(sum alpha_k): the
    normaliser of tf.contrib.distributions.Dirichlet.log_prob (dirichlet_mix.py:111-113)."""
    return gammaln(alpha).sum(0) - gammaln(alpha.sum(0))


def dirichlet_prior(class_counts, class_prior='data'):
    """dirichlet_mix.py:116-129."""
    class_counts = np.asarray(class_counts, np.float32)
    uniform_prior = 1.0 / 14
    data_prior = (class_counts / (1e-20 + class_counts.sum())).astype('float32')
    if class_prior == 'uniform':
        return np.float32(uniform_prior)
    if class_prior == 'data':
        return data_prior
    weight = float(class_prior)
    prior = weight * uniform_prior + (1 - weight) * data_prior
    return (prior / prior.sum())


def dirichlet_fusion(probs, dirichlet_params, prior, sigma=1.0, dtype=np.float32):
    """dirichlet_mix.py:14-36 with the graph side :100-113.

    probs: list of [N,H,W,C] softmax outputs (renormalised here as :100-102 does);
    dirichlet_params: list of [C_out, C_gt] arrays (column c = concentration of gt class c);
    returns the fused score [N,H,W,C_gt] (argmax over the last axis is the prediction)."""
    total = None
    for p, params in zip(probs, dirichlet_params):
        p = p.astype(dtype)
        p = p / p.sum(axis=-1, keepdims=True)
        alpha = (np.asarray(sigma, dtype) * params.astype('float32')).astype(dtype)
        logx = np.log(np.asarray(1e-20, dtype) + p)                   # [N,H,W,C_out]
        unnorm = logx @ (alpha - np.asarray(1, dtype))                 # sum_k (a_kc-1) log x_k
        ll = unnorm - dirichlet_log_norm(alpha.astype(dtype)).astype(dtype)
        total = ll if total is None else total + ll
    return total + np.log(np.asarray(1e-20, dtype) + np.asarray(prior, dtype))


def _log_f32(x):
    """Correctly rounded float32 logarithm: evaluate in float64, round once."""
    with np.errstate(divide='ignore'):
        return np.log(x.astype(np.float64)).astype(np.float32)


def dirichlet_fusion_f32(probs, dirichlet_params, prior, sigma=1.0):
    """dirichlet_mix.py:14-36,100-113 in float32 with a FIXED operation order - the definition
    the device's exact mode reproduces bit for bit (TensorFlow's own reduction order is not
    observable offline, so the order is fixed here and documented in DESIGN.md):
      s      = p_0 + p_1 + ... + p_{C-1}                  sequential, one rounding per add
      x_k    = 1e-20 + p_k / s                            IEEE division, then the add
      L_k    = RN(log x_k)                                correctly rounded float32 logarithm
      u_c    = (..((a_0c L_0) + (a_1c L_1)) + ...)        a_kc = sigma*alpha_kc - 1; every product
                                                          and every sum rounded (no fused FMA)
      ll_c   = u_c - lbeta(sigma*alpha[:, c])
      score  = ((ll^0 + ll^1) + ...) + log(1e-20 + prior) experts in order, prior last
    All arrays float32; returns the fused score [..., C]."""
    f32 = np.float32
    total = None
    for p, params in zip(probs, dirichlet_params):
        p = np.asarray(p, f32)
        c_out = p.shape[-1]
        alpha = f32(sigma) * np.asarray(params).astype(f32)
        am1 = (alpha - f32(1)).astype(f32)
        log_norm = dirichlet_log_norm(alpha.astype(np.float64)).astype(f32)
        s = p[..., 0].copy()
        for k in range(1, c_out):
            s = s + p[..., k]
        logx = _log_f32(f32(1e-20) + p / s[..., None])
        acc = logx[..., 0:1] * am1[0]
        for k in range(1, c_out):
            acc = acc + logx[..., k:k + 1] * am1[k]
        ll = acc - log_norm
        total = ll if total is None else total + ll
    with np.errstate(divide='ignore'):
        log_prior = np.log(f32(1e-20) + np.broadcast_to(np.asarray(prior, f32),
                                                         (total.shape[-1],))).astype(f32)
    assert total.dtype == f32
    return total + log_prior


def dirichlet_uncertainty_fusion(probs, conditional_params, uncertainties, prior,
                                 dtype=np.float64):
    """uncertainty_dirichlet_mix.py:18-52: per pixel alpha = cond*(1-mix) + mix*(I+1) with
    mix = mean_c(var) / max(var over the whole tensor); score[c] = sum_m Dirichlet(alpha[:,c])
    .log_prob(1e-20 + p_m) + log(1e-20 + prior[c])."""
    num_classes = probs[0].shape[-1]
    standard = (np.eye(num_classes) + np.ones((num_classes, num_classes))).astype(dtype)
    total = None
    for p, cond, unc in zip(probs, conditional_params, uncertainties):
        unc = unc.astype(dtype)
        mix = (unc.mean(axis=-1) / unc.max())[..., None, None]
        alpha = cond.astype(dtype) * (1 - mix) + mix * standard          # [..., C_out, C_gt]
        logx = np.log(np.asarray(1e-20, dtype) + p.astype(dtype))[..., None]
        ll = ((alpha - 1) * logx).sum(-2) - (gammaln(alpha).sum(-2) - gammaln(alpha.sum(-2)))
        total = ll if total is None else total + ll
    return total + np.log(np.asarray(1e-20, dtype) + np.asarray(prior, dtype))


# ------------------------------------------------------------ average / variance / MC
def average_fusion(probs):
    """average_mix.py:18-21: argmax(mean_m prob_m)."""
    return np.mean(np.stack(probs), axis=0)


def variance_fusion(probs, variances):
    """variance_mix.py:7-15: inverse-variance weighted mean of the experts' probabilities;
    variances are [N,H,W,1] (mean over classes, keepdims) as built at variance_mix.py:65-66."""
    dtype = probs[0].dtype
    certainties = np.stack([1 / (np.asarray(1e-20, dtype) + v) for v in variances], axis=0)
    probs = np.stack(probs, axis=0)
    return (certainties * probs).sum(0) / certainties.sum(0)


def mc_moments(samples, axis=0):
    """tf.nn.moments over the sample axis: population mean / variance
    (variance_mix.py:65, bayesian_fcn.py:55)."""
    mean = samples.mean(axis=axis, keepdims=True)
    var = ((samples - mean) ** 2).mean(axis=axis)
    return mean.squeeze(axis), var


def normed_entropy(x, axis=-1):
    """custom_layers.py:251-256."""
    return -(x * np.log(np.clip(x, 1e-10, 1.0))).sum(axis=axis) / np.log(
        np.asarray(x.shape[axis], x.dtype))


def sampling_uncertainty(samples):
    """bayesian_fcn.py:48-57; samples [T,N,H,W,C]."""
    mean, var = mc_moments(samples, 0)
    return mean, {'entropy': normed_entropy(mean),
                  'cond_entropy': normed_entropy(samples).mean(axis=0),
                  'variance': var.sum(axis=-1)}


def sufficient_statistics(prob, labels, num_classes):
    """dirichlet_mix.py:142-163: S[c,k] = sum_{label==c} log(1e-10 + prob[...,k]) and
    n[c] = #{label==c}; float64 accumulation like the host loop :187-203."""
    flat_p = prob.reshape(-1, prob.shape[-1])
    flat_l = labels.reshape(-1)
    logp = np.log(np.asarray(1e-10, prob.dtype) + flat_p).astype(np.float64)
    stats = np.zeros((num_classes, prob.shape[-1]))
    counts = np.zeros(num_classes, np.int64)
    for c in range(num_classes):
        sel = flat_l == c
        stats[c] = logp[sel].sum(0)
        counts[c] = sel.sum()
    return stats, counts
