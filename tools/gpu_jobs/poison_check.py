"""Uninitialised-memory hunt: fill the caching allocator's free blocks with NaN, then run the inference pass, the training
forward and the backward; any kernel that reads memory it (or a predecessor) never wrote turns its output into NaN."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
import bench as B  # noqa: E402
from bench_train import clean_batch  # noqa: E402
from diffusesg_b200.model.diffusesg.train_graph import train_state  # noqa: E402
from diffusesg_b200.utils.synthetic import CONFIGS  # noqa: E402

dev = torch.device("cuda:0")
cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "vg"]
Bt = int(sys.argv[2]) if len(sys.argv) > 2 else 128


def poison(gb=24):
    junk = [torch.full((1 << 28,), float("nan"), device=dev) for _ in range(gb)]   # 1 GiB each
    junk += [torch.full((1 << 30,), float("nan"), device=dev) for _ in range(6)]    # 4 GiB each: the workspace-sized blocks
    small = [torch.full((n,), float("nan"), device=dev) for n in (1 << 10, 1 << 14, 1 << 18, 1 << 22) for _ in range(64)]
    del junk, small
    torch.cuda.synchronize()


torch.manual_seed(0)
np.random.seed(0)
model = B.build_native_model(cfg, dev).train()
net = model.model
adj, node, flags = [t.to(dev) for t in clean_batch(cfg, Bt, 1234)]
sig = (torch.randn(Bt, device=dev) * 1.2 - 1.2).exp()
na = adj + torch.randn_like(adj) * sig.view(-1, 1, 1, 1)
nx = node + torch.randn_like(node) * sig.view(-1, 1, 1)
f4 = flags[:, None, :, None] & flags[:, None, None, :]
na, nx = na * f4, nx * flags[:, :, None]


def nans(*ts):
    return [int(torch.isnan(t).sum()) for t in ts]


poison()
with torch.no_grad():
    sa, sn = net.denoise(na, nx, flags, sig, None, None)
torch.cuda.synchronize()
print("inference pass (n_cond = B, no self-cond): NaNs", nans(sa, sn))
poison()
with torch.no_grad():
    sa2, sn2 = net.denoise(na, nx, flags, sig, sa.nan_to_num(), sn.nan_to_num())
print("inference pass with self-cond: NaNs", nans(sa2, sn2))
poison()
da, dn = net.denoise(na, nx, flags, sig, None, None)
torch.cuda.synchronize()
print("training forward, no self-cond: NaNs", nans(da, dn), "requires_grad", da.requires_grad)
poison()
da2, dn2 = net.denoise(na, nx, flags, sig, sa.nan_to_num(), sn.nan_to_num())
print("training forward with self-cond: NaNs", nans(da2, dn2))
poison()
(da2.square().mean() + dn2.square().mean()).backward()
torch.cuda.synchronize()
ts = train_state(net, dev)
bad = [k for k in ts.order if torch.isnan(ts.g(k)).any()]
print("backward: tensors with NaN gradients:", len(bad), bad[:12])
