"""Timeline of CTA 0 of the first fused-MLP launch of width DSG_TRACE_C (default 192) of a denoiser pass."""
import os
import sys

os.environ.setdefault("DSG_TRACE_C", "192")
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_native_model  # noqa: E402
from diffusesg_b200 import native  # noqa: E402
from diffusesg_b200.utils.synthetic import CONFIGS, synthetic_inputs  # noqa: E402

cfg = CONFIGS["vg"]
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
G0 = int(sys.argv[2]) if len(sys.argv) > 2 else 6
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
print("MMA (17): fc1 ready / last W1 landed | fc2 H ready / last W2 landed")
print("workers (0..15): ready, acc1 available, GELU stored, arrived | output: start / acc2 complete / issued")
for g in range(G0, G0 + 14):
    m = rel[g, 17, :].tolist()
    w = rel[g, :16, :]
    def rng(e):
        v = w[:, e][w[:, e] >= 0]
        return f"{int(v.min())}-{int(v.max())}" if len(v) else "-"
    print(f"g={g:2d} MMA fc1 {m[0]}/{m[1]} fc2 {m[2]}/{m[3]} | W {rng(0)} {rng(1)} {rng(2)} {rng(4)} | out {rng(5)} {rng(6)} {rng(7)}")
