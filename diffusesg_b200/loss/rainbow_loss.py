"""EDM regression loss for node + adjacency attributes (loss/rainbow_loss.py:5-99 of the reference).

The masked, weighted squared-error reduction over the [B, C_e, N, N] / [B, N, C_n] tensors is one native launch
(dsg_edm_loss_sums) and, when the predictions require grad, an autograd node whose backward is one more
(dsg_edm_loss_sums_backward): ``loss.backward()`` works against any model that produced the predictions with an
autograd graph (the reference's PyTorch denoiser; the native denoiser has no backward pass yet).  The
[B]-sized normalisation keeps the reference's expressions, including its use of ``edge_loss_weight`` for the node term
under reduction='mean' (:84-85).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from diffusesg_b200 import native


class NodeAdjRainbowLoss(nn.Module):
    def __init__(self, edge_loss_weight, node_loss_weight, objective, flag_reweight=False):
        super().__init__()
        assert objective in ["score", "diffusion", "edm"], "Loss mode {:s} is not supported!".format(objective)
        self.edge_loss_weight = edge_loss_weight
        self.node_loss_weight = node_loss_weight
        self.flag_reweight = flag_reweight
        self.objective = objective

    def forward(self, net_pred_a, net_pred_x, net_target_a, net_target_x, net_cond, adjs_perturbed, adjs_gt,
                x_perturbed, x_gt, node_flags, loss_weight=None, cond_val=None, flag_matching=False, reduction="mean"):
        if flag_matching:
            raise ValueError("Graph matching is not supported for node-adj loss!")
        return self.get_regression_loss(net_pred_a, net_pred_x, net_target_a, net_target_x, net_cond, node_flags, None,
                                        loss_weight, cond_val, reduction)

    def get_regression_loss(self, pred_adj, pred_node, target_adj, target_node, net_cond, node_flags, reweight_coef,
                            loss_weight, condition_true_values, reduction):
        if self.objective == "score":
            raise NotImplementedError
        if reweight_coef is not None or node_flags.dim() != 2 or pred_adj.dim() != 4 or pred_node.dim() != 3:
            raise NotImplementedError("only [B, C, N, N] / [B, N, F] tensors with [B, N] flags and no reweighting are built")
        if torch.is_grad_enabled() and (pred_adj.requires_grad or pred_node.requires_grad):
            # trainer path (runner/trainer/trainer_node_adj.py:112-173): loss.backward() reaches the predictions through
            # the fused backward kernel; the [B]-sized normalisation below is ordinary torch arithmetic
            s_adj, s_node = native.edm_loss_sums_autograd(pred_adj, target_adj, pred_node, target_node, loss_weight, node_flags)
        else:
            s_adj, s_node = native.edm_loss_sums(pred_adj, target_adj, pred_node, target_node, loss_weight, node_flags)
        num_node_entries = node_flags.sum(dim=-1)      # [B]   (:80-82)
        num_adj_entries = num_node_entries ** 2
        if reduction == "mean":
            loss_adj = s_adj.sum() / num_adj_entries * self.edge_loss_weight
            loss_node = s_node.sum() / num_node_entries * self.edge_loss_weight
        elif reduction is None or reduction == "none":
            loss_adj = s_adj / num_adj_entries / pred_adj.size(1) * self.edge_loss_weight
            loss_node = s_node / num_node_entries / pred_node.size(-1) * self.node_loss_weight
        else:
            raise NotImplementedError(f"reduction={reduction!r}")
        return loss_adj, loss_node
