"""Run a few preconditioned denoiser passes (VG geometry) - the unit ncu profiles.

    python tools/one_pass.py [--batch 512] [--passes 2] [--config vg]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_native_model  # noqa: E402
from diffusesg_b200.utils.synthetic import CONFIGS, synthetic_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--config", default="vg")
ap.add_argument("--dense", action="store_true", help="no padding skipping (the sampler's default is to skip)")
args = ap.parse_args()
cfg = CONFIGS[args.config]
dev = torch.device("cuda:0")
model = build_native_model(cfg, dev)
adj, node, flags, sigmas, sc_adj, sc_node = [t.to(dev) for t in synthetic_inputs(cfg, args.batch, seed=7)]
sig = torch.tensor(1.5, device=dev).view(-1).expand(args.batch)
import contextlib  # noqa: E402
plan = None if args.dense else model.model.make_skip_plan(flags)
skipping = model.model.skipping(plan) if plan is not None else contextlib.nullcontext()
skipping.__enter__()
print("padding skipping:", "off" if plan is None else f"kept {plan.kept_fraction:.3f} of the stage-0 pixels, buckets {plan.counts} x {plan.sides}, level 2: {plan.level2[1:] if plan.level2 else None}")
with torch.no_grad():
    for _ in range(args.passes):
        a, n = model.model.denoise(adj, node, flags, sig, sc_adj, sc_node)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
with torch.no_grad():
    a, n = model.model.denoise(adj, node, flags, sig, sc_adj, sc_node)
e1.record()
torch.cuda.synchronize()
print("pass ms", e0.elapsed_time(e1), "finite", bool(torch.isfinite(a).all() and torch.isfinite(n).all()))
