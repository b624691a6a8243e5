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


class NodeAdjEDMObjectiveGenerator:
    """Training objective generator for node + adjacency attributes (runner/objectives/edm.py:130-281 of the
    reference: EDMObjectiveGenerator / NodeAdjEDMObjectiveGenerator), scene-graph flavour: precond = sigma_dist =
    'edm', non-symmetric noise, [B, N] node flags.

    Same constructor, method names, return tuples and RNG consumption order (sigma draw, adjacency noise, node
    noise, all on ``dev``).  The noising itself is one fused native launch (dsg_train_noise) whose outputs are
    bit-identical to the reference's fp32 expressions; the [B]-sized sigma / weight / coefficient math stays torch.
    """

    def __init__(self, precond, sigma_dist, *, other_params=None, dev=None, objective="edm", symmetric_noise=True):
        if precond != "edm" or sigma_dist != "edm":
            raise NotImplementedError("only precond = sigma_dist = 'edm' is built (config/edm_diffuse_sg/*.yaml)")
        if symmetric_noise:
            raise NotImplementedError("symmetric_noise=True is not built: the scene-graph configs pass False")
        self.precond, self.sigma_dist = precond, sigma_dist
        self.objective, self.dev = objective, dev
        self.symmetric_noise = symmetric_noise
        self.edm_params = get_edm_params()
        self.other_params = other_params

    def get_training_sigmas_weights(self, num_samples):
        """:158-176, 'edm' branch."""
        rnd_normal = torch.randn(num_samples, device=self.dev)
        sigmas = (rnd_normal * self.edm_params.P_std + self.edm_params.P_mean).exp()
        weights = (sigmas ** 2 + self.edm_params.sigma_data ** 2) / (sigmas * self.edm_params.sigma_data) ** 2
        return sigmas, weights

    def get_network_input(self, clean_adjs, clean_x=None, node_flags=None, sigmas=None, *args, **kwargs):
        """:233-254 -> (noisy_adjs, noise_added_to_adjs, noisy_x, noise_added_to_x)."""
        from diffusesg_b200 import native
        assert len(sigmas) == len(clean_adjs)
        if node_flags.dim() != 2 or clean_adjs.dim() != 4 or clean_x.dim() != 3:
            raise NotImplementedError("only [B, C, N, N] adjacency / [B, N, F] node tensors with [B, N] flags are built")
        eps_adj = torch.randn_like(clean_adjs)   # reference order: adjacency noise first (graph_utils.py:140)
        eps_x = torch.randn_like(clean_x)        # then node noise (edm.py:245)
        return native.train_noise(clean_adjs, clean_x, eps_adj, eps_x, sigmas, node_flags)

    def get_input_output(self, clean_adjs, clean_x=None, node_flags=None, *args, **kwargs):
        """:256-281 -> (net_input_a, net_input_x, net_cond, net_target_a, net_target_x, (c_skip, c_out, c_in, c_noise,
        sigmas, weights))."""
        sigmas, weights = self.get_training_sigmas_weights(clean_adjs.size(0))
        c_skip, c_out, c_in, c_noise = get_preconditioning_params(self.precond, sigmas, None, None, self.edm_params)
        noisy_adjs, _, noisy_x, _ = self.get_network_input(clean_adjs, clean_x, node_flags, sigmas)
        return noisy_adjs, noisy_x, sigmas, clean_adjs, clean_x, (c_skip, c_out, c_in, c_noise, sigmas, weights)
