"""torchrun --nproc-per-node 2: per-step loss / gradient norm of the data-parallel training step on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.getcwd())
import bench as B  # noqa: E402
from bench_train import EMA_COEFS, LR, MAX_NORM, clean_batch  # noqa: E402
from diffusesg_b200.loss.rainbow_loss import NodeAdjRainbowLoss  # noqa: E402
from diffusesg_b200.runner.objectives.edm import NodeAdjEDMObjectiveGenerator  # noqa: E402
from diffusesg_b200.runner.trainer.trainer_node_adj import train_one_step  # noqa: E402
from diffusesg_b200.utils.synthetic import CONFIGS  # noqa: E402
from diffusesg_b200.utils.train_utils import FusedAdam, NativeDDP, NativeEMA  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
force = os.environ.get("FORCE_DDP") == "1"
if world > 1 or force:
    if force and world == 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29655")
        dist.init_process_group("nccl", device_id=dev, rank=0, world_size=1)
    else:
        dist.init_process_group("nccl", device_id=dev)
cfg = CONFIGS["vg"]
torch.manual_seed(1234 + rank)
np.random.seed(1234 + rank)
model = B.build_native_model(cfg, dev).train()
emas = [NativeEMA(model, beta=c) for c in EMA_COEFS] if os.environ.get("NO_EMA") != "1" else []
opt = FusedAdam(model, lr=LR, max_grad_norm=MAX_NORM)
opt.attach_emas(emas)
wrapped = NativeDDP(model) if (world > 1 or force) else model
gen = NodeAdjEDMObjectiveGenerator("edm", "edm", dev=dev, symmetric_noise=False)
loss_fn = NodeAdjRainbowLoss(1.0, 1.0, "edm")
adj, node, flags = [t.to(dev) for t in clean_batch(cfg, int(os.environ.get("BATCH", 128)), 1234 + rank)]
if os.environ.get("STAGES") == "1":
    from diffusesg_b200 import native
    net = model.model
    na, nx, cond, ta, tx, (c_skip, c_out, c_in, c_noise, sigmas, weights) = gen.get_input_output(adj, node, flags)
    if os.environ.get("POISON_WS") == "1":      # hand the first call a workspace full of a chosen bit pattern
        pat = int(os.environ.get("PATTERN", "0x7fc00000"), 16)
        junk = [torch.full((1 << 29,), pat, dtype=torch.int32, device=dev) for _ in range(8)]
        del junk
    for n_st in [0, 0, 1, -1]:
        native.lib().dsg_debug_set_stop_after(n_st)
        with torch.no_grad():
            sa, sn = net.denoise(na, nx, flags, sigmas, None, None)
        torch.cuda.synchronize()
        nat = net.__dict__["_nat"]
        x = nat.debug_buffer("X", torch.float32)
        y = nat.debug_buffer("Y", torch.bfloat16)
        if n_st == 0:
            xv = x[: adj.shape[0] * 4096 * 96].view(adj.shape[0], 64, 64, 96)
            bad = torch.isnan(xv).any(-1)                     # [B, i, j]
            per = bad.flatten(1).sum(1)
            nb = flags.sum(1)
            who = per.nonzero().flatten()[:6].tolist()
            print(f"rank {rank}: bad pixels per sample (first): {[(b, int(per[b]), int(nb[b])) for b in who]}; "
                  f"rows with bad pixels in sample {who[0] if who else None}: "
                  f"{bad[who[0]].any(1).nonzero().flatten().tolist()[:20] if who else None} cols "
                  f"{bad[who[0]].any(0).nonzero().flatten().tolist()[:20] if who else None}; rc NaNs "
                  f"{int(torch.isnan(nat.debug_buffer('rc', torch.float32)).sum())} coef NaNs "
                  f"{int(torch.isnan(nat.debug_buffer('coef', torch.float32)).sum())}", flush=True)
        print(f"rank {rank} stages {n_st}: X NaNs {int(torch.isnan(x).sum())} of {x.numel()} Y NaNs {int(torch.isnan(y.float()).sum())} "
              f"film NaNs {int(torch.isnan(nat.debug_buffer('film', torch.float32)).sum())} out NaNs {int(torch.isnan(sa).sum())}", flush=True)
    native.lib().dsg_debug_set_stop_after(-1)
if os.environ.get("TRACE") == "1":
    _net = model.model
    _orig = _net.denoise

    def traced(adjs, nodes, node_flags, sigmas, sa=None, sn=None):
        out = _orig(adjs, nodes, node_flags, sigmas, sa, sn)
        torch.cuda.synchronize()
        print(f"rank {rank} trace: denoise grad={torch.is_grad_enabled()} in NaNs {int(torch.isnan(adjs).sum())} "
              f"{int(torch.isnan(nodes).sum())} sigma [{float(sigmas.min()):.4g}, {float(sigmas.max()):.4g}] sc "
              f"{None if sa is None else int(torch.isnan(sa).sum())} out NaNs {int(torch.isnan(out[0]).sum())} "
              f"{int(torch.isnan(out[1]).sum())} |out| {float(out[0].detach().nan_to_num().norm()):.4f} "
              f"nan_w {int(torch.isnan(opt.ts.flat).sum())}", flush=True)
        return out
    _net.denoise = traced
if os.environ.get("PRECHECK") == "1":
    sig = torch.full((adj.shape[0],), 0.7, device=dev)
    net = model.model
    print(f"rank {rank} pre: nan_w {int(torch.isnan(opt.ts.flat).sum())} |w| {float(opt.ts.flat.norm()):.4f}", flush=True)
    for rep in range(3):
        with torch.no_grad():
            sa, sn = net.denoise(adj, node, flags, sig, None, None)
        torch.cuda.synchronize()
        print(f"rank {rank} pre: inference pass {rep}: NaNs {int(torch.isnan(sa).sum())} {int(torch.isnan(sn).sum())} "
              f"|out| {float(sa.nan_to_num().norm()):.4f}", flush=True)
    da, dn = net.denoise(adj, node, flags, sig, None, None)
    torch.cuda.synchronize()
    print(f"rank {rank} pre: training forward: NaNs {int(torch.isnan(da).sum())} {int(torch.isnan(dn).sum())} |out| {float(da.nan_to_num().norm()):.4f}", flush=True)
    del da, dn
for it in range(int(os.environ.get("ITERS", 12))):
    p0 = model.raw_passes
    la, ln = train_one_step(wrapped, opt, emas or None, gen, loss_fn, adj, node, flags, MAX_NORM)
    torch.cuda.synchronize()
    ts = opt.ts
    lo = ts.offs["read_out.0.weight"]
    print(f"rank {rank} it {it}: loss {float(la.mean() + ln.mean()):.5f} passes {model.raw_passes - p0} gnorm {float(opt.grad_norm()):.4f} "
          f"|g_rest| {float(ts.grad[:lo].norm()):.4f} |g_heads| {float(ts.grad[lo:].norm()):.4f} |w| {float(ts.flat.norm()):.4f} "
          f"nan_g {int(torch.isnan(ts.grad).sum())} nan_w {int(torch.isnan(ts.flat).sum())}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
