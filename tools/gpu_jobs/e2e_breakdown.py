"""Where the host-side time of one end-to-end sample() call goes (VG, batch 512)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from bench import build_native_model, make_sampler  # noqa: E402
from diffusesg_b200.utils.synthetic import CONFIGS, synthetic_node_flags  # noqa: E402

cfg = CONFIGS["vg"]
dev = torch.device("cuda:0")
B = 512
print("cpu threads", torch.get_num_threads(), "cores", os.cpu_count())
for k in range(3):
    t = time.perf_counter(); a = torch.randn(B, 6, 64, 64); t1 = time.perf_counter() - t
    buf = torch.empty(B, 6, 64, 64, pin_memory=True)
    t = time.perf_counter(); buf.normal_(); t2 = time.perf_counter() - t
    t = time.perf_counter(); d = buf.to(dev, non_blocking=True); torch.cuda.synchronize(); t3 = time.perf_counter() - t
    t = time.perf_counter(); h = torch.empty(d.shape, pin_memory=True); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); t4 = time.perf_counter() - t
    t = time.perf_counter(); c = d.cpu(); t5 = time.perf_counter() - t
    print(f"randn {t1:.3f}  pinned normal_ {t2:.3f}  h2d {t3:.4f}  d2h pinned(+alloc) {t4:.4f}  d2h .cpu() {t5:.4f}")
model = build_native_model(cfg, dev)
s = make_sampler(cfg, dev, 256)
flags = synthetic_node_flags(cfg, B, seed=1234)
torch.manual_seed(0); np.random.seed(0)
for k in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    a, n = s.sample(model=model, node_flags=flags, num_node_chan=12, num_edge_chan=6)
    t_e2e = time.perf_counter() - t
    ia, inn = torch.randn(B, 6, 64, 64, device=dev), torch.randn(B, 64, 12, device=dev)
    fd = flags.to(dev)
    torch.cuda.synchronize(); t = time.perf_counter()
    s.sample_on_device(model, fd, init_adjs=ia, init_nodes=inn, num_node_chan=12, num_edge_chan=6)
    t_host = time.perf_counter() - t
    torch.cuda.synchronize(); t_dev = time.perf_counter() - t
    print(f"e2e {t_e2e:.3f} s   device-resident {t_dev:.3f} s (host returned after {t_host:.3f} s)")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
a, n = s.sample(model=model, node_flags=flags, num_node_chan=12, num_edge_chan=6)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
