"""EDM constants, noise schedule and preconditioning coefficients (host side).

Same names and values as runner/objectives/edm.py:60-129 of the reference.  Only the 'edm' variant is built:
it is the only one the shipped DiffuseSG configs select (config/edm_diffuse_sg/*.yaml: mcmc.precond = edm).
"""
from __future__ import annotations

from collections import namedtuple

import torch

EDM_PARAMS = namedtuple("EDM_PARAMS", ["sigma_min_training", "sigma_max_training", "sigma_min_sampling",
                                       "sigma_max_sampling", "sigma_data", "P_mean", "P_std", "rho"])


def get_edm_params():
    """runner/objectives/edm.py:60-63."""
    return EDM_PARAMS(sigma_min_training=0.0, sigma_max_training=float("inf"), sigma_min_sampling=0.002,
                      sigma_max_sampling=80.0, sigma_data=0.5, P_mean=-1.2, P_std=1.2, rho=7)


def get_edm_sigma_from_t(t):
    return torch.as_tensor(t)


def get_edm_sigma_deriv_t(t):
    return torch.ones_like(torch.as_tensor(t))


def get_edm_t_from_sigma(sigma):
    return torch.as_tensor(sigma)


def get_preconditioning_params(precond, sigmas, vp_params=None, ve_params=None, edm_params=None):
    """c_skip, c_out, c_in, c_noise (runner/objectives/edm.py:111-129, 'edm' branch)."""
    if precond != "edm":
        raise NotImplementedError(f"precond={precond!r}: only 'edm' is built (the DiffuseSG configs use nothing else)")
    p = edm_params or get_edm_params()
    c_skip = p.sigma_data ** 2 / (sigmas ** 2 + p.sigma_data ** 2)
    c_out = sigmas * p.sigma_data / (sigmas ** 2 + p.sigma_data ** 2).sqrt()
    c_in = 1 / (p.sigma_data ** 2 + sigmas ** 2).sqrt()
    c_noise = sigmas.log() / 4
    return c_skip, c_out, c_in, c_noise
