"""Timeline of CTA 0 of the first fused-MLP launch of a denoiser pass (clock64 stamps), printed per chunk."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_native_model  # noqa: E402
from diffusesg_b200 import native  # noqa: E402
from diffusesg_b200.utils.synthetic import CONFIGS, synthetic_inputs  # noqa: E402

cfg = CONFIGS["vg"]
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
model = build_native_model(cfg, dev)
adj, node, flags, sigmas, sc_adj, sc_node = [t.to(dev) for t in synthetic_inputs(cfg, B, seed=7)]
sig = torch.tensor(1.5, device=dev).view(-1).expand(B)
with torch.no_grad():
    for _ in range(2):
        model.model.denoise(adj, node, flags, sig, sc_adj, sc_node)
    buf = torch.zeros(64 * 18 * 8, dtype=torch.int64, device=dev)
    native.lib().dsg_debug_trace_next_mlp(buf.data_ptr())
    model.model.denoise(adj, node, flags, sig, sc_adj, sc_node)
torch.cuda.synchronize()
t = buf.cpu().view(64, 18, 8)
t0 = int(t[t > 0].min())
rel = torch.where(t > 0, t - t0, torch.full_like(t, -1))
print("MMA warp (17): ev0 fc1 ready, ev1 fc1 last W1 landed, ev2 fc2 H ready, ev3 fc2 last W2 landed")
print("workers (0..15): ev0 ready, ev1 acc1 available, ev2 before h_empty, ev3 sH free, ev4 chunk done, ev5/6/7 output")
for g in range(12):
    mma = rel[g, 17, :8].tolist()
    w = rel[g, :16, :]
    def rng(e):
        v = w[:, e][w[:, e] >= 0]
        return f"{int(v.min())}-{int(v.max())}" if len(v) else "-"
    print(f"g={g:2d} MMA {mma}  W ready {rng(0)} acc1 {rng(1)} pre-h {rng(2)} hfree {rng(3)} done {rng(4)} out {rng(5)}/{rng(6)}/{rng(7)}")
