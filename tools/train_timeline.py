"""Per-kernel time of one training step (CUPTI through torch.profiler): which kernels the step spends its time in.

    python tools/train_timeline.py [--config vg] [--batch 128] [--out profiles/...txt]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
import bench as B  # noqa: E402
from bench_train import EMA_COEFS, LR, MAX_NORM, clean_batch  # noqa: E402
from diffusesg_b200.loss.rainbow_loss import NodeAdjRainbowLoss  # noqa: E402
from diffusesg_b200.runner.objectives.edm import NodeAdjEDMObjectiveGenerator  # noqa: E402
from diffusesg_b200.runner.trainer.trainer_node_adj import train_one_step  # noqa: E402
from diffusesg_b200.utils.synthetic import CONFIGS  # noqa: E402
from diffusesg_b200.utils.train_utils import FusedAdam, NativeEMA  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="vg")
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--out", default=None)
args = ap.parse_args()
dev = torch.device("cuda:0")
cfg = CONFIGS[args.config]
torch.manual_seed(0)
np.random.seed(0)
model = B.build_native_model(cfg, dev).train()
emas = [NativeEMA(model, beta=c) for c in EMA_COEFS]
opt = FusedAdam(model, lr=LR, max_grad_norm=MAX_NORM)
opt.attach_emas(emas)
gen = NodeAdjEDMObjectiveGenerator("edm", "edm", dev=dev, symmetric_noise=False)
loss_fn = NodeAdjRainbowLoss(edge_loss_weight=1.0, node_loss_weight=1.0, objective="edm")
adj, node, flags = [t.to(dev) for t in clean_batch(cfg, args.batch, 1)]
for _ in range(3):
    train_one_step(model, opt, emas, gen, loss_fn, adj, node, flags)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
p0 = model.raw_passes
e0.record()
for _ in range(args.steps):
    train_one_step(model, opt, emas, gen, loss_fn, adj, node, flags)
e1.record()
torch.cuda.synchronize()
head = (f"{args.config} batch {args.batch}: {e0.elapsed_time(e1) / args.steps:.2f} ms per training step (unprofiled, "
        f"{(model.raw_passes - p0) / args.steps:.1f} forward passes per step); peak memory "
        f"{torch.cuda.max_memory_allocated() / 2**30:.1f} GiB\n")
from torch.profiler import ProfilerActivity, profile  # noqa: E402
np.random.seed(5)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(args.steps):
        train_one_step(model, opt, emas, gen, loss_fn, adj, node, flags)
    torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:80]
        r = rows.setdefault(name, [0, 0.0])
        r[0] += 1
        r[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
total = sum(r[1] for r in rows.values())
text = head + f"profiled: {total / 1e3 / args.steps:.2f} ms of kernel time per step, {sum(r[0] for r in rows.values()) // args.steps} launches per step\n"
for name, (cnt, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:40]:
    text += f"{us / 1e3 / args.steps:9.3f} ms {100 * us / total:5.1f} %  x{cnt // args.steps:<5d} {name}\n"
print(text)
if args.out:
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").write(text)
