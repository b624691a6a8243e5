"""torchrun --nproc-per-node 2: the data-parallel training step over NCCL against the same step on one GPU.

Each rank runs forward + backward on its half of a batch through NativeDDP (gradient all-reduce inside the tape's backward);
rank 0 then repeats the whole batch alone.  The loss is a mean over the batch, so the averaged DDP gradient must equal the
single-GPU gradient (up to the bf16 GEMM rounding of different batch tilings)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.getcwd())
import bench as B  # noqa: E402
from bench_train import clean_batch  # noqa: E402
from diffusesg_b200.loss.rainbow_loss import NodeAdjRainbowLoss  # noqa: E402
from diffusesg_b200.model.diffusesg.train_graph import train_state  # noqa: E402
from diffusesg_b200.utils.synthetic import CONFIGS  # noqa: E402
from diffusesg_b200.utils.train_utils import NativeDDP, find_denoiser  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "vg"]
Bt = 8
torch.manual_seed(7 + rank)      # different initial weights: the wrapper must broadcast rank 0's
model = B.build_native_model(cfg, dev).train()
model.self_condition = False     # no coin flip: both runs take the same path
with torch.no_grad():
    for p in model.parameters():
        p.add_(torch.randn_like(p) * 1e-3)
adj, node, flags = [t.to(dev) for t in clean_batch(cfg, Bt, 5)]
g = torch.Generator().manual_seed(3)
sig = (torch.randn(Bt, generator=g) * 1.2 - 1.2).exp().to(dev)
noisy_a = adj + torch.randn(adj.shape, generator=g).to(dev) * sig.view(-1, 1, 1, 1)
noisy_x = node + torch.randn(node.shape, generator=g).to(dev) * sig.view(-1, 1, 1)
f4 = flags[:, None, :, None] & flags[:, None, None, :]
noisy_a, noisy_x = noisy_a * f4, noisy_x * flags[:, :, None]
loss_fn = NodeAdjRainbowLoss(1.0, 1.0, "edm")


def run(m, sl):
    for p in m.parameters():
        p.grad = None
    oa, ox = m(adjs=noisy_a[sl], nodes=noisy_x[sl], node_flags=flags[sl], sigmas=sig[sl])
    la, ln = loss_fn(oa, ox, adj[sl], node[sl], None, None, None, None, None, flags[sl], loss_weight=None, reduction="none")
    (la.mean() + ln.mean()).backward()
    torch.cuda.synchronize()


ddp = NativeDDP(model)
ts = train_state(find_denoiser(model), dev)
half = Bt // world
run(ddp, slice(rank * half, (rank + 1) * half))
g_ddp = ts.grad.clone()
both = [torch.empty_like(g_ddp) for _ in range(world)]
dist.all_gather(both, g_ddp)
same = all(torch.equal(both[0], b) for b in both)
if rank == 0:
    ts.ddp_group = None
    run(model, slice(0, Bt))
    rel = float((g_ddp - ts.grad).norm() / ts.grad.norm())
    print(f"ddp_check {cfg['dataset']}: ranks hold identical averaged gradients: {same}; "
          f"rel-L2 vs the single-GPU gradient of the whole batch: {rel:.3e}")
    assert same and rel < 2e-2
dist.barrier()
dist.destroy_process_group()
