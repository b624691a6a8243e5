"""Sample-sharded generation across the GPUs of one box.

The reference shards the conditioning set with a DistributedSampler and a per-GPU batch of
``batch_size // world_size`` (runner/sampler/sampler_utils.py:32-36), offsets the RNG seed by the rank
(utils/arg_parser.py:293-294) and gathers the generated tensors once at the end
(runner/sampler/sampler_node_adj.py:331-345).  This module is that data path without the dataset / metric
machinery around it: each rank samples its own slice with zero per-step communication.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
from torch import distributed as dist

from ...utils.dist_training import gather_tensors


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of ``total`` samples owned by ``rank`` (sizes differ by <= 1)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def per_gpu_batch(batch_size: int, world: int) -> int:
    """runner/sampler/sampler_utils.py:34."""
    return max(1, batch_size // world)


def seed_everything(seed: int, rank: int) -> None:
    """utils/arg_parser.py:293-299: every rank draws different noise and different self-conditioning coins."""
    torch.manual_seed(seed + rank)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed + rank)
    np.random.seed(seed + rank)


def sample_sharded(sampler, model, node_flags_all: torch.Tensor, batch_size: int, num_node_chan: int,
                   num_edge_chan: int, gather_device=None):
    """Generate one graph per row of ``node_flags_all`` [S, N] (identical on every rank).

    Each rank runs ``sampler.sample`` over its slice in chunks of ``batch_size // world`` graphs, then the
    slices are all-gathered in rank order (padded to equal length for the collective and trimmed afterwards).
    Returns CPU tensors (adjs [S, C_e, N, N], nodes [S, N, C_n]) on every rank.  Without an initialised
    process group this is plain single-GPU batched sampling.
    """
    ddp = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size() if ddp else 1
    rank = dist.get_rank() if ddp else 0
    total = node_flags_all.shape[0]
    begin, end = shard_bounds(total, world, rank)
    step = per_gpu_batch(batch_size, world)
    adjs, nodes = [], []
    for lo in range(begin, end, step):
        a, n = sampler.sample(model=model, node_flags=node_flags_all[lo:min(end, lo + step)],
                              num_node_chan=num_node_chan, num_edge_chan=num_edge_chan)
        adjs.append(a)
        nodes.append(n)
    n_img = node_flags_all.shape[1]
    adjs = (adjs[0] if len(adjs) == 1 else torch.cat(adjs)) if adjs else torch.zeros(0, num_edge_chan, n_img, n_img)
    nodes = (nodes[0] if len(nodes) == 1 else torch.cat(nodes)) if nodes else torch.zeros(0, n_img, num_node_chan)
    if not ddp:
        return adjs, nodes
    longest = -(-total // world)
    dev = gather_device if gather_device is not None else getattr(sampler, "dev", "cpu")

    def pad(t):
        if t.shape[0] == longest:
            return t
        return torch.cat([t, t.new_zeros((longest - t.shape[0],) + tuple(t.shape[1:]))])

    def to_host(t):
        if not t.is_cuda:
            return t
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)   # pinned: the gathered set is world x the local one
        h.copy_(t, non_blocking=True)
        return h

    ga = to_host(gather_tensors(pad(adjs), 0, dev))
    gn = to_host(gather_tensors(pad(nodes), 0, dev))
    if ga.is_pinned() and torch.cuda.is_available():
        torch.cuda.current_stream().synchronize()
    if total % world == 0:
        return ga, gn            # equal shards: nothing was padded
    keep = torch.cat([torch.arange(r * longest, r * longest + (shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0]))
                      for r in range(world)])
    return ga[keep], gn[keep]
