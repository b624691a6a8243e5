"""CPU oracle for the DiffuseSG denoiser forward (TEST INFRASTRUCTURE ONLY).

This is a plain fp32 torch restatement of the reference's algorithm, written
functionally over a ``state_dict`` so that it carries no module tree of its
own.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product path in
``diffusesg_b200/`` never does (it fails loudly when the CUDA library is
missing).

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this file
against golden vectors produced by executing the unmodified reference modules
(``tests/golden/make_golden.py``, run in the build container where
``/root/reference`` exists).  The reference's own test-suite holds no vectors
for this path.

Every function cites the reference lines it restates; paths are relative to
``/root/reference/DiffuseSG/``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- #
# static geometry
# --------------------------------------------------------------------------- #
def stage_plan(img: int, embed: int, depths: Sequence[int], heads: Sequence[int], window: int):
    """Enumerate the Swin blocks of the U-Net in execution order.

    Restates the constructor bookkeeping of model/diffusesg/diffusesg.py:656-702
    (stage dims/resolutions), :189-192 (window clamp, shift cancel) and :459
    (odd blocks shift by window//2).
    """
    n = len(depths)
    plan = []
    for s in range(n):
        res = img // (2 ** s)
        for j in range(depths[s]):
            plan.append(dict(prefix=f"down_layers.{s}.blocks.{j}", dim=embed * 2 ** s, res=res,
                             heads=heads[s], **_win(res, window, j)))
    for u in range(n):
        s = n - 1 - u
        res = img // (2 ** s)
        for j in range(depths[s]):
            plan.append(dict(prefix=f"up_layers.{u}.blocks.{j}", dim=embed * 2 ** s, res=res,
                             heads=heads[s], **_win(res, window, j)))
    return plan


def _win(res: int, window: int, j: int):
    shift = 0 if j % 2 == 0 else window // 2
    if res <= window:
        return dict(window=res, shift=0)
    return dict(window=window, shift=shift)


def relative_position_index(w: int) -> Tensor:
    """[w*w, w*w] index into the (2w-1)^2 bias table (diffusesg.py:88-97)."""
    ys, xs = torch.meshgrid(torch.arange(w), torch.arange(w), indexing="ij")
    py, px = ys.reshape(-1), xs.reshape(-1)
    dy = py[:, None] - py[None, :] + (w - 1)
    dx = px[:, None] - px[None, :] + (w - 1)
    return dy * (2 * w - 1) + dx


def shift_attn_mask(res: int, w: int, s: int) -> Tensor:
    """[nW, w*w, w*w] additive mask, 0 / -100 (diffusesg.py:207-226)."""
    def region(v):
        # 0: [0, res-w)   1: [res-w, res-s)   2: [res-s, res)
        return (v >= res - w).long() + (v >= res - s).long()
    r = region(torch.arange(res))
    ids = (r[:, None] * 3 + r[None, :]).float()                        # [res, res]
    nw = res // w
    ids = ids.view(nw, w, nw, w).permute(0, 2, 1, 3).reshape(nw * nw, w * w)
    diff = ids[:, None, :] - ids[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


# --------------------------------------------------------------------------- #
# building blocks
# --------------------------------------------------------------------------- #
def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _ln(sd, name, x):
    w = sd[name + ".weight"]
    return F.layer_norm(x, (w.numel(),), w, sd[name + ".bias"], 1e-5)


def film_silu(sd, prefix, x, emb):
    """silu(shift + x * (1 + scale)) with (scale, shift) = affine(emb).chunk(2)
    (diffusesg.py:238-240 and :574-576)."""
    p = _lin(sd, prefix + ".affine", emb)[:, None, :]
    c = x.shape[-1]
    scale, shift = p[..., :c], p[..., c:]
    return F.silu(torch.addcmul(shift, x, scale + 1))


def noise_embedding(sd, noise_labels: Tensor, embed: int) -> Tensor:
    """Sinusoid(embed) -> Linear -> SiLU -> Linear -> SiLU
    (diffusesg.py:507-513, :768-771)."""
    half = embed // 2
    freqs = torch.arange(half, dtype=torch.float32, device=noise_labels.device) / half
    freqs = (1.0 / 10000.0) ** freqs
    ang = noise_labels.float().ger(freqs)
    e = torch.cat([ang.cos(), ang.sin()], dim=1)
    e = F.silu(_lin(sd, "map_layer0", e))
    return F.silu(_lin(sd, "map_layer1", e))


def window_attention(sd, prefix, xw, heads, w, mask):
    """xw [B*nW, T, C] -> same (diffusesg.py:108-139)."""
    bw, t, c = xw.shape
    d = c // heads
    qkv = _lin(sd, prefix + ".qkv", xw).view(bw, t, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * d ** -0.5, qkv[1], qkv[2]
    att = q @ k.transpose(-1, -2)                                        # [bw, h, T, T]
    idx = relative_position_index(w).to(xw.device)
    bias = sd[prefix + ".relative_position_bias_table"][idx.reshape(-1)].view(t, t, heads).permute(2, 0, 1)
    att = att + bias[None]
    if mask is not None:
        nw = mask.shape[0]
        att = (att.view(bw // nw, nw, heads, t, t) + mask[None, :, None]).view(bw, heads, t, t)
    att = att.softmax(dim=-1)
    out = (att @ v).transpose(1, 2).reshape(bw, t, c)
    return _lin(sd, prefix + ".proj", out)


def swin_block(sd, blk, x, emb, capture=None):
    """One SwinTransformerBlock (diffusesg.py:232-277).  x [B, L, C]."""
    p, res, w, s, heads = blk["prefix"], blk["res"], blk["window"], blk["shift"], blk["heads"]
    b, l, c = x.shape
    x = film_silu(sd, p, x, emb)                     # the shortcut is the modulated tensor (:242)
    y = _ln(sd, p + ".norm1", x).view(b, res, res, c)
    if s > 0:
        y = torch.roll(y, shifts=(-s, -s), dims=(1, 2))
    nw = res // w
    yw = y.view(b, nw, w, nw, w, c).permute(0, 1, 3, 2, 4, 5).reshape(b * nw * nw, w * w, c)
    mask = shift_attn_mask(res, w, s).to(x.device) if s > 0 else None
    aw = window_attention(sd, p + ".attn", yw, heads, w, mask)
    a = aw.view(b, nw, nw, w, w, c).permute(0, 1, 3, 2, 4, 5).reshape(b, res, res, c)
    if s > 0:
        a = torch.roll(a, shifts=(s, s), dims=(1, 2))
    x = x + a.view(b, l, c)
    h = _lin(sd, p + ".mlp.fc1", _ln(sd, p + ".norm2", x))
    x = x + _lin(sd, p + ".mlp.fc2", F.gelu(h))
    if capture is not None:
        capture[p] = x
    return x


def patch_merging(sd, prefix, x, res):
    """2x2 space-to-depth, LN(4C), Linear 4C->2C (diffusesg.py:314-335)."""
    b, l, c = x.shape
    g = x.view(b, res // 2, 2, res // 2, 2, c)       # [b, y, dy, x, dx, c]
    # channel blocks in the order (dy,dx) = (0,0),(1,0),(0,1),(1,1)
    g = torch.cat([g[:, :, 0, :, 0], g[:, :, 1, :, 0], g[:, :, 0, :, 1], g[:, :, 1, :, 1]], dim=-1)
    g = g.reshape(b, l // 4, 4 * c)
    return F.linear(_ln(sd, prefix + ".norm", g), sd[prefix + ".reduction.weight"])


def patch_breakup(sd, prefix, x, res):
    """Linear -> LN -> depth-to-space -> LN -> Linear (diffusesg.py:374-403)."""
    b, l, d = x.shape
    y = _ln(sd, prefix + ".norm", F.linear(x, sd[prefix + ".pre_linear.weight"]))
    co = d // 4
    y = y.view(b, res, res, 4, co)
    out = y.new_zeros(b, res, 2, res, 2, co)         # [b, y, dy, x, dx, c]
    out[:, :, 0, :, 0] = y[:, :, :, 0]
    out[:, :, 1, :, 0] = y[:, :, :, 1]
    out[:, :, 0, :, 1] = y[:, :, :, 2]
    out[:, :, 1, :, 1] = y[:, :, :, 3]
    out = out.reshape(b, 4 * l, co)
    return F.linear(_ln(sd, prefix + ".post_norm", out), sd[prefix + ".post_linear.weight"])


def mask_pairs(t: Tensor, flags: Tensor) -> Tensor:
    """Zero rows and columns of padded nodes of [B, C, N, N] (utils/graph_utils.py:5-38)."""
    bad = ~flags.bool()
    return t.masked_fill(bad[:, None, :, None], 0.0).masked_fill(bad[:, None, None, :], 0.0)


def mask_rows(t: Tensor, flags: Tensor) -> Tensor:
    """Zero padded rows of [B, N, C] (utils/graph_utils.py:41-86)."""
    return t.masked_fill(~flags.bool()[:, :, None], 0.0)


# --------------------------------------------------------------------------- #
# full forward
# --------------------------------------------------------------------------- #
def denoiser_forward(sd: Dict[str, Tensor], *, img: int, embed: int, depths: Sequence[int],
                     heads: Sequence[int], window: int, self_condition: bool,
                     adj: Tensor, node: Tensor, flags: Tensor, noise_labels: Tensor,
                     sc_adj: Optional[Tensor] = None, sc_node: Optional[Tensor] = None,
                     capture: Optional[dict] = None):
    """Raw network F(adj, node | flags, c_noise) -> (adj_out [B,Ce,N,N], node_out [B,N,Cn]).

    Restates DiffuseSG.forward / forward_features (diffusesg.py:739-830) for the
    scene-graph case (multi-channel adj, [B,N] flags, symmetric_noise=False).
    """
    b, n = flags.shape
    emb = noise_embedding(sd, noise_labels, embed)
    node_t = node.float().permute(0, 2, 1)                                # [B, Cn, N]
    if self_condition:                                                       # self-cond FIRST (:791-794)
        adj = torch.cat([torch.zeros_like(adj) if sc_adj is None else sc_adj, adj], dim=1)
        sc_t = torch.zeros_like(node_t) if sc_node is None else sc_node.float().permute(0, 2, 1)
        node_t = torch.cat([sc_t, node_t], dim=1)
    rows = node_t[:, :, :, None].expand(-1, -1, -1, n)                      # value of node i at (i, j)
    cols = node_t[:, :, None, :].expand(-1, -1, n, -1)                      # value of node j at (i, j)
    pair_in = torch.cat([adj, mask_pairs(torch.cat([rows, cols], dim=1), flags)], dim=1)

    # patch embed: 1x1 conv, LN, FiLM (:569-577)
    wp = sd["patch_embed.proj.weight"]
    x = F.conv2d(pair_in, wp, sd["patch_embed.proj.bias"]).flatten(2).transpose(1, 2)
    x = _ln(sd, "patch_embed.norm", x)
    x = film_silu(sd, "patch_embed", x, emb)
    if capture is not None:
        capture["patch_embed"] = x

    nl = len(depths)
    plan = stage_plan(img, embed, depths, heads, window)
    it = iter(plan)
    skips: List[Tensor] = []
    for s in range(nl):                                                      # encoder (:746-748)
        for _ in range(depths[s]):
            x = swin_block(sd, next(it), x, emb, capture)
        if s < nl - 1:
            x = patch_merging(sd, f"down_layers.{s}.downsample", x, img // 2 ** s)
            if capture is not None:
                capture[f"down_layers.{s}.downsample"] = x
        skips.append(x)
    for u in range(nl):                                                      # decoder (:751-756)
        s = nl - 1 - u
        skip = skips.pop()
        if u > 0:                                                            # first decoder stage drops its skip
            x = patch_breakup(sd, f"up_layers.{u}.upsample", torch.cat([x, skip], dim=-1), img // 2 ** (s + 1))
            if capture is not None:
                capture[f"up_layers.{u}.upsample"] = x
        for _ in range(depths[s]):
            x = swin_block(sd, next(it), x, emb, capture)

    x = _ln(sd, "norm", x)                                                   # (:758)
    rep = x.view(b, n, n, embed).permute(0, 3, 1, 2)
    rep = F.conv_transpose2d(rep, sd["read_out.0.weight"], sd["read_out.0.bias"])
    rep = F.conv2d(rep, sd["read_out.1.weight"], sd["read_out.1.bias"])
    rep = F.conv2d(rep, sd["read_out.2.weight"], sd["read_out.2.bias"])    # [B, 96, N, N]
    if capture is not None:
        capture["shared_rep"] = rep

    def head(prefix, t):
        return _lin(sd, prefix + ".fc2", F.gelu(_lin(sd, prefix + ".fc1", t)))

    adj_out = head("readout_adj_mlp", rep.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
    pooled = mask_pairs(rep, flags).mean(dim=-1)                              # sum over j / N (:812-813)
    node_out = head("readout_node_mlp", pooled.permute(0, 2, 1))
    return mask_pairs(adj_out, flags), mask_rows(node_out, flags)


# --------------------------------------------------------------------------- #
# EDM preconditioning wrapper
# --------------------------------------------------------------------------- #
SIGMA_DATA = 0.5


def precond_coefficients(sigmas: Tensor):
    """c_skip, c_out, c_in, c_noise for 'edm' (runner/objectives/edm.py:122-126)."""
    sd2 = SIGMA_DATA ** 2
    c_skip = sd2 / (sigmas ** 2 + sd2)
    c_out = sigmas * SIGMA_DATA / (sigmas ** 2 + sd2).sqrt()
    c_in = 1 / (sd2 + sigmas ** 2).sqrt()
    c_noise = sigmas.log() / 4
    return c_skip, c_out, c_in, c_noise


def precond_forward(net, adjs, nodes, flags, sigmas, sc_adjs=None, sc_nodes=None, coin=None):
    """D(x; sigma) = mask(c_skip x + c_out F(c_in x, c_noise, self_cond))
    (model/precond/precond.py:65-110).

    ``net(adj, node, flags, c_noise, sc_adj, sc_node)`` is the raw network.
    ``coin`` is a zero-argument callable returning the uniform draw that the
    reference takes from ``np.random.rand()`` at precond.py:90; a draw < 0.5
    triggers the extra pass that refreshes the self-conditioning tensors.
    """
    c_skip, c_out, c_in, c_noise = precond_coefficients(sigmas)
    ca = lambda c: c.view(-1, 1, 1, 1)
    cn = lambda c: c.view(-1, 1, 1)

    def run(sa, sn):
        fa, fn = net(ca(c_in) * adjs, cn(c_in) * nodes, flags, c_noise, sa, sn)
        da = mask_pairs(ca(c_skip) * adjs + ca(c_out) * fa.float(), flags)
        dn = mask_rows(cn(c_skip) * nodes + cn(c_out) * fn.float(), flags)
        return da, dn

    if coin is not None and coin() < 0.5:
        sc_adjs, sc_nodes = run(sc_adjs, sc_nodes)
    return run(sc_adjs, sc_nodes)


def dataset_channels(name: str):
    """(C_e, C_n, allowed_nodes) for the 'bits' encoding with boxes
    (utils/sg_utils.py:348-409)."""
    if "visual_genome" in name:
        nt, et, allowed = 150, 51, 62
    elif "coco_stuff" in name:
        nt, et, allowed = 171, 7, 33
    else:
        raise NotImplementedError(name)
    return math.ceil(math.log2(et)), math.ceil(math.log2(nt)) + 4, allowed
