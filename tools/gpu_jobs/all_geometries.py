"""A short eager sampler run per geometry (padding skipping at both levels, fused decode): a quick all-geometry check, and
the unit to put under a memory checker where one is available (compute-sanitizer is closed on this pool)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from bench import build_native_model, make_sampler  # noqa: E402
from diffusesg_b200.utils.synthetic import CONFIGS, synthetic_node_flags  # noqa: E402

dev = torch.device("cuda:0")
for name, batch in (("tiny", 6), ("vg", 6), ("coco", 5)):
    cfg = CONFIGS[name]
    model = build_native_model(cfg, dev)
    s = make_sampler(cfg, dev, 2)
    s.use_graphs = False
    flags = synthetic_node_flags(cfg, batch, seed=3)
    torch.manual_seed(0)
    np.random.seed(0)
    out = s.sample_decoded(model, flags, 7, 150, num_node_chan=cfg["c_n"], num_edge_chan=cfg["c_e"], return_state=True)
    print(name, "ok", [tuple(t.shape) for t in out], bool(torch.isfinite(out[0]).all()))
