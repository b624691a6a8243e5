"""Warm per-launch timeline of one denoiser pass: CUDA events around every launch (library-side brackets).

    python tools/pass_timeline.py [--batch 512] [--config vg] [--csv gpurun_out/timeline.csv]

Prints launch order with ms, achieved TFLOP/s and GB/s (algorithmic flops / bytes), then totals per label.
"""
import argparse
import collections
import csv
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_native_model  # noqa: E402
from diffusesg_b200 import native  # noqa: E402
from diffusesg_b200.utils.synthetic import CONFIGS, synthetic_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--config", default="vg")
ap.add_argument("--dense", action="store_true", help="no padding skipping (the sampler's default is to skip)")
ap.add_argument("--csv", default="gpurun_out/timeline.csv")
ap.add_argument("--quiet", action="store_true")
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
    for _ in range(3):
        model.model.denoise(adj, node, flags, sig, sc_adj, sc_node)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        model.model.denoise(adj, node, flags, sig, sc_adj, sc_node)
    e1.record()
    torch.cuda.synchronize()
    print(f"pass (unbracketed, mean of 3): {e0.elapsed_time(e1) / 3:.3f} ms")
    native.profile_begin(1)
    model.model.denoise(adj, node, flags, sig, sc_adj, sc_node)
    torch.cuda.synchronize()
    os.makedirs(os.path.dirname(args.csv) or ".", exist_ok=True)
    native.profile_dump(args.csv)
    native.profile_read()
    native.profile_stop()
rows = list(csv.DictReader(open(args.csv)))
tot = sum(float(r["ms"]) for r in rows)
print(f"{len(rows)} launches, {tot:.3f} ms bracketed")
agg = collections.OrderedDict()
for r in rows:
    ms = float(r["ms"])
    tf = float(r["flops"]) / ms / 1e9 if ms > 0 else 0
    gb = float(r["bytes"]) / ms / 1e6 if ms > 0 else 0
    if not args.quiet:
        print(f"{int(r['index']):3d} {r['label']:34s} rows {int(r['rows']):9d} {ms * 1e3:8.1f} us {tf:7.1f} TF/s {gb:7.0f} GB/s")
    k = (r["label"], r["rows"])
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += ms; a[2] += float(r["flops"]); a[3] += float(r["bytes"])
print("--- by label")
sys.stdout.flush()
for (label, nrows), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{label:34s} rows {nrows:>9s} x{a[0]:2d} {a[1] * 1e3:8.1f} us {100 * a[1] / tot:5.1f}%  {a[2] / a[1] / 1e9:7.1f} TF/s {a[3] / a[1] / 1e6:7.0f} GB/s")
