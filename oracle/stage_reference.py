"""Stage the UNMODIFIED reference hot-path modules for timing on the GPU box (TEST / BENCH INFRASTRUCTURE ONLY).

The reference is pure Python, so there is nothing to compile: "building" ``oracle/_ref`` means copying the files
the path consists of, byte for byte, from where they lie under ``/root/reference`` into ``oracle/_ref/DiffuseSG/``
(git-ignored: reference sources never enter this repository's history; NOT gpurun-ignored: the directory travels
to the GPU box next to the built ``.so``, where ``/root/reference`` does not exist).  A manifest with the SHA-256 of
every staged file is written next to them; ``load()`` verifies it before importing.

    python -m oracle.stage_reference          # stage (needs /root/reference), same as __graft_entry__.build()

``load()`` imports the staged (or, in the build container, the original) modules with the three-name stand-in for
``timm.models.layers`` the image lacks (DropPath at rate 0 is the identity, to_2tuple, trunc_normal_ ==
torch.nn.init.trunc_normal_) and returns the reference's own classes.

Only tests/, __graft_entry__ and bench.py's reference / cpu_baseline legs may import this module.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/DiffuseSG"
DST = os.path.join(HERE, "_ref", "DiffuseSG")
MANIFEST = os.path.join(HERE, "_ref", "MANIFEST.json")

# the files the sampling hot path (and its training-objective neighbour) consists of
FILES = [
    "model/diffusesg/diffusesg.py",        # a6-a15: the Swin-UNet denoiser
    "model/precond/precond.py",            # a5: EDM preconditioning + self-conditioning coin flip
    "runner/mcmc_sampler/__init__.py",     # GeneralSampler base
    "runner/mcmc_sampler/edm.py",          # a1-a3: the stochastic Heun loop
    "runner/objectives/__init__.py",
    "runner/objectives/edm.py",            # a4: schedule / preconditioning parameters, a17 objective
    "utils/graph_utils.py",                # a16: mask_adjs / mask_nodes
    "utils/attribute_code.py",             # bin2dec (decode rule, f-1)
    "loss/rainbow_loss.py",                # a17 loss
]


def _sha(path: str) -> str:
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(verbose: bool = True) -> bool:
    """Copy FILES from /root/reference (if present).  Returns True when oracle/_ref is complete afterwards."""
    if os.path.isdir(SRC):
        manifest = {}
        for rel in FILES:
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(SRC, rel), dst)
            manifest[rel] = _sha(dst)
        with open(MANIFEST, "w") as f:
            json.dump({"source": SRC, "files": manifest}, f, indent=1)
        if verbose:
            print(f"oracle/_ref: staged {len(FILES)} unmodified reference files")
    return staged()


def staged() -> bool:
    if not os.path.exists(MANIFEST):
        return False
    files = json.load(open(MANIFEST))["files"]
    return all(os.path.exists(os.path.join(DST, rel)) and _sha(os.path.join(DST, rel)) == h for rel, h in files.items()) \
        and set(files) == set(FILES)


def root() -> str | None:
    """Directory to import the reference from: the staged copy, else the original, else None."""
    if staged():
        return DST
    if os.path.isdir(SRC):
        return SRC
    return None


def _timm_stand_in():
    import torch
    layers = types.ModuleType("timm.models.layers")

    class DropPath(torch.nn.Identity):
        def __init__(self, p=0.0):
            super().__init__()

    layers.DropPath = DropPath
    layers.to_2tuple = lambda v: v if isinstance(v, tuple) else (v, v)
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    for name, mod in (("timm", types.ModuleType("timm")), ("timm.models", types.ModuleType("timm.models")),
                      ("timm.models.layers", layers)):
        sys.modules.setdefault(name, mod)


def load():
    """-> namespace with the reference's DiffuseSG, NodeAdjPrecond, NodeAdjEDMSampler, mask_adjs, mask_nodes, bin2dec,
    NodeAdjEDMObjectiveGenerator, NodeAdjRainbowLoss and ``where`` (the directory they were imported from)."""
    base = root()
    if base is None:
        raise RuntimeError("the reference is neither staged under oracle/_ref (run __graft_entry__.build() in the "
                           "build container) nor present at /root/reference")
    _timm_stand_in()
    sys.path.insert(0, base)
    try:
        from model.diffusesg.diffusesg import DiffuseSG
        from model.precond.precond import NodeAdjPrecond
        from runner.mcmc_sampler.edm import NodeAdjEDMSampler
        from runner.objectives.edm import NodeAdjEDMObjectiveGenerator
        from loss.rainbow_loss import NodeAdjRainbowLoss
        from utils.attribute_code import bin2dec
        from utils.graph_utils import mask_adjs, mask_nodes
    finally:
        sys.path.remove(base)
    return types.SimpleNamespace(DiffuseSG=DiffuseSG, NodeAdjPrecond=NodeAdjPrecond, NodeAdjEDMSampler=NodeAdjEDMSampler,
                                 NodeAdjEDMObjectiveGenerator=NodeAdjEDMObjectiveGenerator,
                                 NodeAdjRainbowLoss=NodeAdjRainbowLoss, bin2dec=bin2dec, mask_adjs=mask_adjs,
                                 mask_nodes=mask_nodes, where=base)


def build_network(ref, cfg, state_dict=None):
    """The reference's own ``NodeAdjPrecond(DiffuseSG(...))`` with get_network's keyword values
    (utils/learning_utils.py:47-74), optionally loaded with ``state_dict`` (strict)."""
    net = ref.DiffuseSG(img_size=cfg["img"], in_chans=cfg["c_e"] + 2 * cfg["c_n"], patch_size=1, embed_dim=cfg["embed"],
                        depths=cfg["depths"], num_heads=[3, 6, 12, 24], window_size=cfg["window"], mlp_ratio=4.,
                        drop_rate=0., attn_drop_rate=0., drop_path_rate=0.0, self_condition=cfg["self_cond"],
                        symmetric_noise=False, out_chans_adj=cfg["c_e"], out_chans_node=cfg["c_n"])
    if state_dict is not None:
        net.load_state_dict(state_dict, strict=True)
    return ref.NodeAdjPrecond(precond="edm", model=net.eval(), self_condition=cfg["self_cond"], symmetric_noise=False).eval()


if __name__ == "__main__":
    ok = stage()
    print("staged:", ok, "->", DST if ok else "(reference not available)")
