"""Two eager training iterations (VG, batch 128 by default): the unit to put under ncu (`-k regex:<kernel> -c <n>`)."""
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
for _ in range(args.steps):
    la, ln = train_one_step(model, opt, emas, gen, loss_fn, adj, node, flags)
torch.cuda.synchronize()
print("loss", float(la.mean() + ln.mean()))
