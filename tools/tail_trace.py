"""Timeline of CTA 0 of the first fused block-tail launch of a denoiser pass (clock64 stamps), printed per chunk."""
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
G0 = int(sys.argv[2]) if len(sys.argv) > 2 else 12
model = build_native_model(cfg, dev)
adj, node, flags, sigmas, sc_adj, sc_node = [t.to(dev) for t in synthetic_inputs(cfg, B, seed=7)]
sig = torch.tensor(1.5, device=dev).view(-1).expand(B)
with torch.no_grad():
    for _ in range(2):
        model.model.denoise(adj, node, flags, sig, sc_adj, sc_node)
    buf = torch.zeros(64 * 19 * 8, dtype=torch.int64, device=dev)
    native.lib().dsg_debug_trace_next_mlp(buf.data_ptr())
    model.model.denoise(adj, node, flags, sig, sc_adj, sc_node)
torch.cuda.synchronize()
t = buf.cpu().view(64, 19, 8)
t0 = int(t[t > 0].min())
rel = torch.where(t > 0, t - t0, torch.full_like(t, -1))
print("MMA(17): fc1 start/issued | fc2 h_full seen/w2 seen | proj start/ready/issued | y_ready seen")
print("TMA(16): w1 issued, w2 issued, att issued, wp issued   X(18): out_ready seen, store read done")
print("workers: G wait/acc1 avail/gelu done/h arrived | P start/proj seen/xin seen/P done")
for g in range(G0, G0 + 14):
    mma = rel[g, 17, :].tolist()
    tma = rel[g, 16, :4].tolist()
    xw = rel[g, 18, :2].tolist()
    w = rel[g, :16, :]
    def rng(e):
        v = w[:, e][w[:, e] >= 0]
        return f"{int(v.min())}-{int(v.max())}" if len(v) else "-"
    print(f"g={g:2d} MMA fc1 {mma[0]}/{mma[1]} fc2 {mma[2]}/{mma[3]} proj {mma[4]}/{mma[5]}/{mma[6]} y {mma[7]} | TMA {tma} X {xw}"
          f" | G {rng(0)} {rng(1)} {rng(2)} {rng(3)} | P {rng(4)} {rng(5)} {rng(6)} {rng(7)}")
