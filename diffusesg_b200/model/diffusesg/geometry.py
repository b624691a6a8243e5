"""Static window geometry of the Swin-style denoiser (host side, tiny tensors).

These are the two persistent buffers the reference stores in its checkpoints
(``attn.relative_position_index`` and ``attn_mask``) plus
the per-stage block table used to drive the native kernels.
"""
from __future__ import annotations

from typing import List, Sequence

import torch


def relative_position_index(w: int) -> torch.Tensor:
    """int64 [w*w, w*w]: row of the (2w-1)^2 bias table used by the token pair (p, q).

    Same values as the buffer built at model/diffusesg/diffusesg.py:88-98 of the
    reference: ((yp - yq + w-1) * (2w-1)) + (xp - xq + w-1).
    """
    t = torch.arange(w * w)
    y, x = t // w, t % w
    return (y[:, None] - y[None, :] + w - 1) * (2 * w - 1) + (x[:, None] - x[None, :] + w - 1)


def shifted_window_mask(res: int, w: int, shift: int) -> torch.Tensor:
    """fp32 [nW, w*w, w*w] additive attention mask of a cyclically shifted block.

    After rolling the grid by ``-shift`` a window may hold tokens from up to
    four disconnected image regions; pairs from different regions get -100 (not
    -inf), as at model/diffusesg/diffusesg.py:207-226.
    """
    edge = torch.arange(res)
    band = (edge >= res - w).to(torch.int64) + (edge >= res - shift).to(torch.int64)
    region = band[:, None] * 3 + band[None, :]
    nw = res // w
    per_win = region.reshape(nw, w, nw, w).transpose(1, 2).reshape(nw * nw, w * w)
    same = per_win[:, :, None] == per_win[:, None, :]
    return torch.where(same, 0.0, -100.0).to(torch.float32)


def block_table(img: int, embed: int, depths: Sequence[int], heads: Sequence[int], window: int) -> List[dict]:
    """Execution-ordered description of every transformer block of the U-Net.

    dim/res per stage follow model/diffusesg/diffusesg.py:656-702; a stage whose
    resolution does not exceed the window uses one window per sample and never
    shifts (:189-192); odd blocks otherwise shift by window // 2 (:459).
    """
    nl = len(depths)
    table: List[dict] = []
    for side in ("down", "up"):
        for k in range(nl):
            s = k if side == "down" else nl - 1 - k
            res = img // (1 << s)
            for j in range(depths[s]):
                if res <= window:
                    w, sh = res, 0
                else:
                    w, sh = window, (window // 2 if j % 2 else 0)
                table.append(dict(prefix=f"{side}_layers.{k}.blocks.{j}", side=side, layer=k, stage=s,
                                  index=j, dim=embed << s, res=res, heads=heads[s], window=w, shift=sh))
    return table
