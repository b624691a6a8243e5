"""CPU restatement of the EDM training objective and loss (TEST INFRASTRUCTURE ONLY: imported by tests/ and by
nothing the product path runs).

Follows, in fp32 torch on the CPU:
  runner/objectives/edm.py:158-176   get_training_sigmas_weights ('edm')
  runner/objectives/edm.py:233-254   NodeAdjEDMObjectiveGenerator.get_network_input (symmetric_noise=False)
  utils/graph_utils.py:122-152       add_sym_normal_noise(non_symmetric=True) with mask_adjs / mask_nodes (:5-86)
  loss/rainbow_loss.py:60-99         NodeAdjRainbowLoss.get_regression_loss ('edm')
Pinned against outputs of the unmodified reference: tests/golden/train_objective.npz (tests/test_oracle_golden.py).
"""
import torch

P_MEAN, P_STD, SIGMA_DATA = -1.2, 1.2, 0.5


def training_sigmas_weights(rnd_normal):
    sigmas = (rnd_normal * P_STD + P_MEAN).exp()
    weights = (sigmas ** 2 + SIGMA_DATA ** 2) / (sigmas * SIGMA_DATA) ** 2
    return sigmas, weights


def _mask_adjs(a, flags):
    m = (flags[:, None, :, None] & flags[:, None, None, :])
    return torch.where(m, a, torch.zeros_like(a))


def _mask_nodes(x, flags):
    return torch.where(flags[:, :, None], x, torch.zeros_like(x))


def network_input(clean_adj, clean_x, flags, sigmas, eps_adj, eps_x):
    """(noisy_adj, noise_adj, noisy_x, noise_x); eps_* are the N(0,1) draws in the reference's order."""
    s4 = sigmas.view(-1, 1, 1, 1)
    noise_a = eps_adj * s4
    noisy_a = clean_adj * torch.ones_like(s4) + noise_a
    noisy_a, noise_a = _mask_adjs(noisy_a, flags), _mask_adjs(noise_a, flags)
    noise_x = _mask_nodes(eps_x * sigmas.view(-1, 1, 1), flags)
    return noisy_a, noise_a, clean_x + noise_x, noise_x


def regression_loss(pred_adj, pred_node, target_adj, target_node, flags, loss_weight, edge_w, node_w, reduction):
    b = pred_adj.shape[0]
    w = torch.ones(b) if loss_weight is None else loss_weight
    sq_a = _mask_adjs((pred_adj - target_adj) ** 2 * 1.0 * w.view(b, 1, 1, 1), flags)
    sq_n = _mask_nodes((pred_node - target_node) ** 2 * 1.0 * w.view(b, 1, 1), flags)
    n_node = flags.sum(dim=-1)
    n_adj = n_node ** 2
    if reduction == "mean":
        return sq_a.sum() / n_adj * edge_w, sq_n.sum() / n_node * edge_w
    la = sq_a.sum(dim=[-1, -2, -3]) / n_adj / sq_a.size(1) * edge_w
    ln = sq_n.sum(dim=[-1, -2]) / n_node / sq_n.size(-1) * node_w
    return la, ln
