"""Mirror of the fusion rule of xview/models/uncertainty_dirichlet_mix.py.  The class
`UncertaintyMix` of that file cannot be constructed as shipped (pre-refactor BaseModel
signature); its per-pixel rule `dirichlet_uncertainty_fusion`
(uncertainty_dirichlet_mix.py:18-52) is what this module provides, on the device."""
import numpy as np
import torch

from .. import device as dev


def dirichlet_uncertainty_fusion(probs, conditional_params, uncertainties, prior):
    """uncertainty_dirichlet_mix.py:18-52: per pixel and expert the Dirichlet parameters are
    blended towards the uninformative I + 1 by mix = mean_k(uncertainty) / max(uncertainty), the
    experts' Dirichlet log-likelihoods of their probability vectors are summed and
    log(1e-20 + prior) is added.

    probs, uncertainties: lists of float32 CUDA tensors [N,H,W,C]; conditional_params: list of
    [C,C] arrays (numpy or CUDA; column c parametrises the Dirichlet of class c); prior: [C].
    Returns the fused class score [N,H,W,C] (float32 CUDA)."""
    device = probs[0].device
    cond = torch.stack([torch.as_tensor(np.asarray(p.detach().cpu() if isinstance(p, torch.Tensor)
                                                   else p, dtype=np.float32))
                        for p in conditional_params]).to(device)
    prior = np.asarray(prior.detach().cpu() if isinstance(prior, torch.Tensor) else prior,
                       dtype=np.float32)
    log_prior = torch.from_numpy(np.log(np.float32(1e-20) + prior).astype(np.float32)).to(device)
    score, _ = dev.dirichlet_uncertainty_fuse([p.contiguous() for p in probs],
                                              [u.contiguous() for u in uncertainties],
                                              cond.contiguous(), log_prior, want_score=True)
    return score
