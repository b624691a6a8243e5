"""One training iteration of the scene-graph denoiser - the body of ``move_forward_one_epoch``
(runner/trainer/trainer_node_adj.py:95-178 of the reference), same call order: objective generator -> zero_grad ->
preconditioned model (with its self-conditioning coin flip) -> loss (reduction='none') -> mean + mean -> backward ->
clip_grad_norm_(10) -> optimizer.step() -> ema.update().  With ``FusedAdam`` the clipping rides inside the optimiser
launch (``max_grad_norm``); with any other optimiser ``nn.utils.clip_grad_norm_`` is called as in the reference."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...utils.train_utils import FusedAdam


def train_one_step(model, optimizer, ema_helper, train_obj_gen, loss_func, adjs_gt, nodes_gt, node_flags,
                   max_grad_norm: float = 10.0):
    """Returns (reg_loss_adj [B], reg_loss_node [B]) detached, like the per-iteration record of :181-182."""
    dev = train_obj_gen.dev
    adjs_gt, nodes_gt, node_flags = adjs_gt.to(dev), nodes_gt.to(dev), node_flags.to(dev)
    net_input_a, net_input_x, net_cond, net_target_a, net_target_x, (c_skip, c_out, c_in, c_noise, sigmas, weights) = \
        train_obj_gen.get_input_output(adjs_gt, nodes_gt, node_flags)
    optimizer.zero_grad(set_to_none=True)
    net_output_a, net_output_x = model(adjs=net_input_a, nodes=net_input_x, node_flags=node_flags, sigmas=sigmas)
    reg_loss_adj, reg_loss_node = loss_func(net_pred_a=net_output_a, net_pred_x=net_output_x, net_target_a=net_target_a,
                                            net_target_x=net_target_x, net_cond=net_cond, adjs_perturbed=net_input_a,
                                            adjs_gt=adjs_gt, x_perturbed=net_input_x, x_gt=nodes_gt, node_flags=node_flags,
                                            loss_weight=weights, reduction="none")
    loss = reg_loss_adj.mean() + reg_loss_node.mean()
    loss.backward()
    if isinstance(optimizer, FusedAdam):
        optimizer.max_grad_norm = max_grad_norm
    else:
        nn.utils.clip_grad_norm_(model.parameters(), max_norm=max_grad_norm, norm_type=2)
    optimizer.step()
    if ema_helper is not None:
        for ema in ema_helper:
            ema.update()
    return reg_loss_adj.detach(), reg_loss_node.detach()
