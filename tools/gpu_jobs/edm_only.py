"""Three eager sampler steps (VG, batch 512): the launches `ncu -k regex:edm_` captures for the EDM step kernels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from bench import build_native_model, make_sampler  # noqa: E402
from diffusesg_b200.utils.synthetic import CONFIGS, synthetic_node_flags  # noqa: E402

cfg = CONFIGS["vg"]
dev = torch.device("cuda:0")
model = build_native_model(cfg, dev)
s = make_sampler(cfg, dev, 3)
s.use_graphs = False
flags = synthetic_node_flags(cfg, 512, seed=1234).to(dev)
torch.manual_seed(0)
np.random.seed(0)
s.sample_on_device(model, flags, num_node_chan=cfg["c_n"], num_edge_chan=cfg["c_e"])
torch.cuda.synchronize()
