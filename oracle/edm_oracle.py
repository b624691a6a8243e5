"""CPU oracle for the EDM stochastic-Heun sampling loop (TEST INFRASTRUCTURE ONLY).

A restatement of ``NodeAdjEDMSampler`` (runner/mcmc_sampler/edm.py:231-445 of
the reference) for the configuration DiffuseSG actually runs:
``solver='heun', discretization='edm', schedule='linear', scaling='none',
alpha=1, symmetric_noise=False``.  Written as a step function over explicit
state so that tests can replay recorded noise and compare step by step.

Parity status: PINNED against (i) the reference's own known-answer mode
(``sanity_check_gt_*``: the final step must return the ground truth exactly,
SURVEY.md section 4) and (ii) golden trajectories produced by the unmodified
reference sampler (``tests/golden/make_golden.py``).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs
may import this module.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from .denoiser_oracle import mask_pairs, mask_rows

Tensor = torch.Tensor

# constructor defaults of the reference sampler (edm.py:236-245) and the EDM
# constants of runner/objectives/edm.py:60-63
NUM_STEPS, S_CHURN, S_MIN, S_MAX, S_NOISE = 256, 40, 0.05, 50, 1.003
SIGMA_MIN, SIGMA_MAX, RHO = 0.002, 80.0, 7


def sigma_grid(num_steps: int = NUM_STEPS) -> Tensor:
    """fp64 Karras grid (edm.py:70, :84-88)."""
    i = torch.arange(num_steps, dtype=torch.float64)
    a, b = SIGMA_MAX ** (1 / RHO), SIGMA_MIN ** (1 / RHO)
    return (a + i / (num_steps - 1) * (b - a)) ** RHO


def t_steps_fp32(num_steps: int = NUM_STEPS) -> Tensor:
    """sigma^-1(round_sigma(grid)) ++ [0], cast to fp32 (edm.py:318-323)."""
    g = sigma_grid(num_steps)
    return torch.cat([g, torch.zeros_like(g[:1])]).to(torch.float32)


def step_scalars(t_cur: Tensor, t_next: Tensor, num_steps: int):
    """All 0-d fp32 scalars one loop iteration derives from (t_cur, t_next)
    (edm.py:355-356, :361, :369, :384-391, :414).  fp32 tensor arithmetic with
    python-float constants, exactly as the reference evaluates it."""
    gamma = min(S_CHURN / num_steps, np.sqrt(2) - 1) if S_MIN <= t_cur <= S_MAX else 0
    t_hat = torch.as_tensor(t_cur + gamma * t_cur)
    noise_coef = (t_hat ** 2 - t_cur ** 2).clip(min=0).sqrt() * 1 * S_NOISE
    h = t_next - t_hat
    t_prime = t_hat + 1 * h
    return dict(gamma=gamma, t_hat=t_hat, noise_coef=noise_coef, h=h, t_prime=t_prime)


def init_sample(flags: Tensor, c_e: int, c_n: int, normal: Callable):
    """Unit-variance start, adjs drawn first (edm.py:257-289, non-symmetric)."""
    b, n = flags.shape
    adjs = mask_pairs(normal((b, c_e, n, n)), flags)
    nodes = mask_rows(normal((b, n, c_n)), flags)
    return adjs, nodes


def heun_step(model, adjs, nodes, flags, sc_a, sc_n, t_cur, t_next, i, num_steps,
              normal: Callable, self_condition: bool = True, gt=None):
    """One iteration of the hot loop (edm.py:350-434).

    ``model(adjs, nodes, flags, sigmas, sc_a, sc_n) -> (D_a, D_n)`` is the
    preconditioned denoiser.  ``gt=(adjs, nodes)`` selects the reference's
    known-answer mode.  Returns (adjs_next, nodes_next, sc_a, sc_n, aux).
    """
    s = step_scalars(t_cur, t_next, num_steps)
    t_hat, h, t_prime = s["t_hat"], s["h"], s["t_prime"]
    adjs_hat = mask_pairs(adjs + s["noise_coef"] * normal(tuple(adjs.shape)), flags)
    nodes_hat = mask_rows(nodes + s["noise_coef"] * normal(tuple(nodes.shape)), flags)
    sig = t_hat.view(-1).expand(flags.size(0))

    def denoise(sa, sn):
        if gt is not None:
            return gt
        da, dn = model(adjs_hat, nodes_hat, flags, sig, sa, sn)
        return mask_pairs(da, flags), mask_rows(dn, flags)

    d1a, d1n = denoise(sc_a, sc_n)
    inv = 1.0 / t_hat                       # sigma'(t)/sigma(t) with sigma(t)=t
    k_a = mask_pairs(inv * adjs_hat - inv * d1a, flags)
    k_n = mask_rows(inv * nodes_hat - inv * d1n, flags)
    prime_a = adjs_hat + h * k_a
    prime_n = nodes_hat + h * k_n
    if i == num_steps - 1:
        next_a, next_n = prime_a, prime_n
        da, dn = d1a, d1n
    else:
        # NB the 2nd evaluation is at (x_hat, t_hat) again, self-conditioned on D1 (edm.py:400-405)
        if gt is None and self_condition:
            sc_a, sc_n = d1a, d1n
        da, dn = denoise(sc_a, sc_n)
        inv_p = 1.0 / t_prime
        kp_a = inv_p * prime_a - inv_p * da
        kp_n = inv_p * prime_n - inv_p * dn
        next_a = adjs_hat + h * (0.5 * k_a + 0.5 * kp_a)
        next_n = nodes_hat + h * (0.5 * k_n + 0.5 * kp_n)
    next_a, next_n = mask_pairs(next_a, flags), mask_rows(next_n, flags)
    if self_condition:
        sc_a, sc_n = da, dn
    aux = dict(adjs_hat=adjs_hat, nodes_hat=nodes_hat, d1=(d1a, d1n), d2=(da, dn), scalars=s)
    return next_a, next_n, sc_a, sc_n, aux


def sample(model, flags: Tensor, c_e: int, c_n: int, num_steps: int = NUM_STEPS,
           normal: Optional[Callable] = None, self_condition: bool = True, gt=None,
           trace: Optional[list] = None):
    """Full loop; returns final (adjs, nodes) (edm.py:291-445)."""
    if normal is None:
        normal = lambda shape: torch.randn(shape)
    ts = t_steps_fp32(num_steps)
    adjs, nodes = init_sample(flags, c_e, c_n, normal)
    adjs, nodes = adjs * ts[0], nodes * ts[0]
    sc_a = sc_n = None
    for i in range(num_steps):
        adjs, nodes, sc_a, sc_n, aux = heun_step(model, adjs, nodes, flags, sc_a, sc_n, ts[i], ts[i + 1],
                                                 i, num_steps, normal, self_condition, gt)
        if trace is not None:
            trace.append((adjs.clone(), nodes.clone()))
    return adjs, nodes


def decode_bits(t: Tensor, num_classes: int) -> Tensor:
    """clamp -> sign -> bits to int -> clamp (runner/sampler/sampler_node_adj.py:222-285 decode rule,
    utils/attribute_code.py:319-328 bin2dec, MSB first).  t [..., nbits] in [-1, 1] -> int64 [...]."""
    bits = (t.clamp(-1, 1) > 0).long()
    nb = bits.shape[-1]
    weights = 2 ** torch.arange(nb - 1, -1, -1)
    return (bits * weights).sum(-1).clamp(0, num_classes - 1)


def decode_samples(adjs: Tensor, nodes: Tensor, flags: Tensor, num_adj_type: int, num_node_type: int):
    """The reference's post-sampling decode for the 'bits' encodings, op for op
    (runner/sampler/sampler_node_adj.py:199-209 boxes, :222-237 _decode_node, :239-285 _decode_adj).
    Returns float tensors like the reference: q_adj [B,N,N], q_node [B,N], bbox [B,N,4]."""
    boxes = mask_rows(nodes[..., -4:] * 0.5 + 0.5, flags)
    node_bits = nodes[..., :-4].clamp(-1.0, 1.0)
    node_bits = torch.where(node_bits > 0.0, torch.ones_like(node_bits), -torch.ones_like(node_bits))
    node_bits = mask_rows(node_bits, flags)
    qb = mask_rows(node_bits.gt(0.0).float(), flags)
    nb = qb.shape[-1]
    weights = 2 ** torch.arange(nb - 1, -1, -1).to(qb.dtype)
    q_node = (weights * qb).sum(-1)
    q_node = q_node.masked_fill(~flags.bool(), 0.0).clamp(min=0, max=num_node_type - 1)
    a = adjs.clamp(-1.0, 1.0)
    a = torch.where(a > 0.0, torch.ones_like(a), -torch.ones_like(a))
    a = mask_pairs(a, flags)
    qa = mask_pairs(a.gt(0.0).float(), flags).permute(0, 2, 3, 1)
    ne = qa.shape[-1]
    q_adj = ((2 ** torch.arange(ne - 1, -1, -1).to(qa.dtype)) * qa).sum(-1)
    pair = flags.bool()[:, :, None] & flags.bool()[:, None, :]
    q_adj = q_adj.masked_fill(~pair, 0.0).clamp(min=0, max=num_adj_type - 1)
    n = flags.shape[1]
    q_adj[:, torch.eye(n).bool()] = 0.0
    return q_adj.contiguous(), q_node, boxes
