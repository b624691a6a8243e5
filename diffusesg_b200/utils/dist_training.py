"""Collectives the sampling path needs (one process per GPU, torch.distributed over NCCL / NVLink).

Sampling shards by sample and needs no per-step communication; the only exchange is the final gather of the
generated graphs (runner/sampler/sampler_node_adj.py:331-345 of the reference).  ``gather_tensors`` has the
reference's signature and semantics (utils/dist_training.py:170-195).
"""
from __future__ import annotations

import torch
from torch import distributed as dist


def gather_tensors(in_tensor: torch.Tensor, cat_dim: int, device) -> torch.Tensor:
    """All-gather ``in_tensor`` from every rank and concatenate along ``cat_dim``, in rank order."""
    world = dist.get_world_size()
    t = in_tensor.to(device).contiguous()
    if cat_dim == 0 and hasattr(dist, "all_gather_into_tensor"):
        shape = list(t.shape)
        shape[0] *= world
        out = torch.empty(shape, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t)
        return out
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return torch.cat(parts, dim=cat_dim)
