"""VE-SDE coefficient functions with the reference's signatures (networks/gf_algorithms/sde.py:15-28,
init_sde :80-116).  Only 've' is on the hot path (`--sde_mode ve`, configs/config.py:34); the other modes
raise.  These run on the host (prior draw on the CPU generator exactly like the reference, sde.py:26-28)."""
import functools

import numpy as np
import torch


def ve_marginal_prob(x, t, sigma_min=0.01, sigma_max=90):
    std = sigma_min * (sigma_max / sigma_min) ** t
    return x, std


def ve_sde(t, sigma_min=0.01, sigma_max=90):
    sigma = sigma_min * (sigma_max / sigma_min) ** t
    drift_coeff = torch.tensor(0)
    diffusion_coeff = sigma * torch.sqrt(torch.tensor(2 * (np.log(sigma_max) - np.log(sigma_min)), device=t.device))
    return drift_coeff, diffusion_coeff


def ve_prior(shape, sigma_min=0.01, sigma_max=90, T=1.0):
    _, sigma_max_prior = ve_marginal_prob(None, T, sigma_min=sigma_min, sigma_max=sigma_max)
    return torch.randn(*shape) * sigma_max_prior


def init_sde(sde_mode):
    """-> prior_fn, marginal_prob_fn, sde_fn, sampling_eps, T   (sde.py:80-116)"""
    if sde_mode != "ve":
        raise NotImplementedError(f"genpose_b200 implements the default VE SDE only (got {sde_mode!r})")
    sigma_min, sigma_max, eps, T = 0.01, 50, 1e-5, 1.0
    marginal_prob_fn = functools.partial(ve_marginal_prob, sigma_min=sigma_min, sigma_max=sigma_max)
    sde_fn = functools.partial(ve_sde, sigma_min=sigma_min, sigma_max=sigma_max)
    prior_fn = functools.partial(ve_prior, sigma_min=sigma_min, sigma_max=sigma_max)
    return prior_fn, marginal_prob_fn, sde_fn, eps, T
