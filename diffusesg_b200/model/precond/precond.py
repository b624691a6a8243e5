"""Drop-in ``NodeAdjPrecond``: EDM preconditioning around the native denoiser.

Interface of model/precond/precond.py:60-114 of the reference: same constructor, same
``forward(adjs, nodes, node_flags, sigmas, self_cond_adjs=None, self_cond_nodes=None) -> (D_adjs, D_nodes)``, same
``round_sigma`` static method, same ``state_dict`` prefix (``model.``), and the same self-conditioning coin flip
drawn from the global numpy RNG (precond.py:90, active in eval mode too), so a run consumes the RNG streams in the
reference's order.  The arithmetic (c_in scaling, c_skip x + c_out F, masks) runs inside the native kernels.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ...runner.objectives.edm import get_edm_params


class NodeAdjPrecond(nn.Module):
    def __init__(self, precond, model, self_condition, symmetric_noise=True):
        super().__init__()
        if precond != "edm":
            raise NotImplementedError(f"precond={precond!r}: only 'edm' is built")
        if symmetric_noise:
            raise NotImplementedError("symmetric_noise=True: scene graphs use non-symmetric noise "
                                      "(utils/learning_utils.py:74)")
        self.precond = precond
        self.model = model
        self.self_condition = self_condition
        self.symmetric_noise = symmetric_noise
        self.edm_params = get_edm_params()
        self.vp_params = self.ve_params = None
        self.raw_passes = 0  # denoiser passes actually executed (the coin flip makes this data dependent)

    def forward(self, adjs, nodes=None, node_flags=None, sigmas=None, self_cond_adjs=None, self_cond_nodes=None,
                *args, **model_kwargs):
        if args or model_kwargs:
            raise NotImplementedError("NodeAdjPrecond (B200): extra model arguments are not supported")
        coin = self.__dict__.get("_forced_coin")   # set by GraphedTrainStep, which draws np.random.rand() itself (same stream)
        if coin is None:
            coin = self.self_condition and np.random.rand() < 0.5
        if self.self_condition and coin:
            with torch.no_grad():
                self_cond_adjs, self_cond_nodes = self.model.denoise(adjs, nodes, node_flags, sigmas,
                                                                     self_cond_adjs, self_cond_nodes)
            self.raw_passes += 1
        self.raw_passes += 1
        return self.model.denoise(adjs, nodes, node_flags, sigmas, self_cond_adjs, self_cond_nodes)

    @staticmethod
    def round_sigma(sigma):
        return torch.as_tensor(sigma)
