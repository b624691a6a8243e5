"""Synthetic scene-graph workloads: configs, node flags and random weights.

The dataset blobs and checkpoints of the reference are not available offline, so
benchmarks and parity tests run on Visual-Genome / COCO-Stuff *shaped* inputs
with seeded random weights.  Shapes follow the reference's configs
(config/edm_diffuse_sg/*.yaml) and channel table (utils/sg_utils.py:348-409).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict

import torch

# name -> network / data geometry.  c_e / c_n are the 'bits' encoding widths
# (ceil(log2(#edge types)), ceil(log2(#node types)) + 4 bbox coordinates).
CONFIGS: Dict[str, dict] = {
    # Visual Genome: N=64, window 8, depths [1,1,3,1]   (…_visual_genome.yaml:8-34)
    "vg": dict(dataset="visual_genome", img=64, window=8, depths=[1, 1, 3, 1], embed=96,
               heads=[3, 6, 12, 24], c_e=6, c_n=12, allowed_nodes=62, self_cond=True),
    # COCO-Stuff: N=40, window 10, depths [1,2,6]        (…_coco.yaml:8-34)
    "coco": dict(dataset="coco_stuff", img=40, window=10, depths=[1, 2, 6], embed=96,
                 heads=[3, 6, 12, 24], c_e=3, c_n=12, allowed_nodes=33, self_cond=True),
    # BASELINE.json config 5: N=64 with 16x16 windows (deeper variant so that shifted 16-windows run)
    "n64w16": dict(dataset="visual_genome", img=64, window=16, depths=[2, 2, 6, 2], embed=96,
                   heads=[3, 6, 12, 24], c_e=6, c_n=12, allowed_nodes=62, self_cond=True),
    # small geometry for fast parity tests: exercises shift, merge, breakup, clamp-to-resolution
    "tiny": dict(dataset="tiny", img=16, window=4, depths=[1, 2, 1], embed=96,
                 heads=[3, 6, 12, 24], c_e=3, c_n=5, allowed_nodes=14, self_cond=True),
}


def in_chans(cfg: dict) -> int:
    """Per-pixel input channels before the self-conditioning doubling
    (utils/sg_utils.py:412-430: in_chans_node + in_chans_adj)."""
    return cfg["c_e"] + 2 * cfg["c_n"]


def synthetic_node_flags(cfg: dict, batch: int, seed: int = 1234) -> torch.Tensor:
    """bool [B, N]; the first n_b ~ U{2..allowed} nodes of each graph are real."""
    g = torch.Generator().manual_seed(seed)
    n_b = torch.randint(2, cfg["allowed_nodes"] + 1, (batch,), generator=g)
    return torch.arange(cfg["img"])[None, :] < n_b[:, None]


def state_dict_spec(cfg: dict) -> "OrderedDict[str, tuple]":
    """Ordered {key: (shape, kind)} of the denoiser's state_dict.

    Mirrors the registration order of the reference module tree
    (model/diffusesg/diffusesg.py:611-720) so that checkpoints interchange.
    ``kind`` drives the synthetic initialiser below.
    """
    embed, depths, heads, img, window = cfg["embed"], cfg["depths"], cfg["heads"], cfg["img"], cfg["window"]
    nl = len(depths)
    cin = in_chans(cfg) * (2 if cfg["self_cond"] else 1)
    spec: "OrderedDict[str, tuple]" = OrderedDict()

    def lin(p, o, i, bias=True):
        spec[p + ".weight"] = ((o, i), "w")
        if bias:
            spec[p + ".bias"] = ((o,), "b")

    def ln(p, c):
        spec[p + ".weight"] = ((c,), "g")
        spec[p + ".bias"] = ((c,), "b")

    def block(p, dim, res, nh, j):
        w, s = (res, 0) if res <= window else (window, 0 if j % 2 == 0 else window // 2)
        if s > 0:
            nw = (res // w) ** 2
            spec[p + ".attn_mask"] = ((nw, w * w, w * w), "mask")
        lin(p + ".affine", 2 * dim, 512)
        ln(p + ".norm1", dim)
        spec[p + ".attn.relative_position_bias_table"] = (((2 * w - 1) ** 2, nh), "t")
        spec[p + ".attn.relative_position_index"] = ((w * w, w * w), "index")
        lin(p + ".attn.qkv", 3 * dim, dim)
        lin(p + ".attn.proj", dim, dim)
        ln(p + ".norm2", dim)
        lin(p + ".mlp.fc1", 4 * dim, dim)
        lin(p + ".mlp.fc2", dim, 4 * dim)

    lin("patch_embed.affine", 2 * embed, 512)
    spec["patch_embed.proj.weight"] = ((embed, cin, 1, 1), "w")
    spec["patch_embed.proj.bias"] = ((embed,), "b")
    ln("patch_embed.norm", embed)
    for s in range(nl):
        dim, res = embed * 2 ** s, img // 2 ** s
        for j in range(depths[s]):
            block(f"down_layers.{s}.blocks.{j}", dim, res, heads[s], j)
        if s < nl - 1:
            lin(f"down_layers.{s}.downsample.reduction", 2 * dim, 4 * dim, bias=False)
            ln(f"down_layers.{s}.downsample.norm", 4 * dim)
    for u in range(nl):
        s = nl - 1 - u
        dim, res = embed * 2 ** s, img // 2 ** s
        if u > 0:
            d = 4 * dim
            lin(f"up_layers.{u}.upsample.pre_linear", d, d, bias=False)
            ln(f"up_layers.{u}.upsample.norm", d)
            lin(f"up_layers.{u}.upsample.post_linear", d // 4, d // 4, bias=False)
            ln(f"up_layers.{u}.upsample.post_norm", d // 4)
        for j in range(depths[s]):
            block(f"up_layers.{u}.blocks.{j}", dim, res, heads[s], j)
    for k in range(3):
        spec[f"read_out.{k}.weight"] = ((embed, embed, 1, 1), "w")
        spec[f"read_out.{k}.bias"] = ((embed,), "b")
    lin("map_layer0", 512, embed)
    lin("map_layer1", 512, 512)
    ln("norm", embed)
    lin("readout_adj_mlp.fc1", embed, embed)
    lin("readout_adj_mlp.fc2", cfg["c_e"], embed)
    lin("readout_node_mlp.fc1", embed, embed)
    lin("readout_node_mlp.fc2", cfg["c_n"], embed)
    return spec


def synthetic_state_dict(cfg: dict, seed: int = 1234, stress: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """Seeded fp32 weights for every key of ``state_dict_spec``.

    ``stress=True`` draws every matrix at ~1/sqrt(fan_in) scale with non-zero
    biases, LayerNorm gains around 1 and a visible relative-position table so
    that each term of the forward contributes numerically (the reference's own
    trunc-normal(0.02) init makes the raw output ~2e-3 rms).
    ``stress=False`` imitates the reference init scale (std 0.02, zero biases).
    Buffers (relative_position_index, attn_mask) are computed, not drawn.
    """
    from ..model.diffusesg.geometry import relative_position_index, shifted_window_mask

    g = torch.Generator().manual_seed(seed)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, (shape, kind) in state_dict_spec(cfg).items():
        if kind == "w":
            fan_in = shape[1] if len(shape) >= 2 else shape[0]
            std = (1.0 / math.sqrt(fan_in)) if stress else 0.02
            out[key] = torch.randn(shape, generator=g) * std
        elif kind == "b":
            out[key] = torch.randn(shape, generator=g) * (0.1 if stress else 0.0)
        elif kind == "g":
            out[key] = 1.0 + torch.randn(shape, generator=g) * (0.1 if stress else 0.0)
        elif kind == "t":
            out[key] = torch.randn(shape, generator=g) * (0.5 if stress else 0.02)
        elif kind == "index":
            w = int(round(math.sqrt(shape[0])))
            out[key] = relative_position_index(w)
        elif kind == "mask":
            nw, t, _ = shape
            w = int(round(math.sqrt(t)))
            res = int(round(math.sqrt(nw))) * w
            out[key] = shifted_window_mask(res, w, w // 2)
        else:  # pragma: no cover
            raise AssertionError(kind)
    return out


def synthetic_inputs(cfg: dict, batch: int, seed: int = 7, sigma: float = 1.5):
    """(adj, node, flags, sigmas, sc_adj, sc_node) fp32 CPU tensors, masked like real sampler state."""
    g = torch.Generator().manual_seed(seed)
    n, ce, cn = cfg["img"], cfg["c_e"], cfg["c_n"]
    flags = synthetic_node_flags(cfg, batch, seed + 1)
    f = flags.float()
    pair = f[:, None, :, None] * f[:, None, None, :]
    adj = torch.randn(batch, ce, n, n, generator=g) * sigma * pair
    node = torch.randn(batch, n, cn, generator=g) * sigma * f[:, :, None]
    sc_adj = torch.randn(batch, ce, n, n, generator=g).clamp(-1.5, 1.5) * pair
    sc_node = torch.randn(batch, n, cn, generator=g).clamp(-1.5, 1.5) * f[:, :, None]
    sigmas = torch.full((batch,), float(sigma))
    return adj, node, flags, sigmas, sc_adj, sc_node
