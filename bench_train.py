#!/usr/bin/env python
"""Benchmark of the DiffuseSG TRAINING step (BASELINE.json config 4): trained scene graphs / second.

    python bench_train.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--config vg] [--batch 128]

A "step" is one iteration of the reference's training loop (runner/trainer/trainer_node_adj.py:95-178): EDM objective
(sigma draw, noising) -> zero_grad -> preconditioned denoiser with its self-conditioning coin flip (a second, no-grad
pass on half of the steps) -> weighted masked loss -> backward -> clip_grad_norm_(10) -> Adam(2e-4) -> the five moving
averages of config/edm_diffuse_sg/*visual_genome.yaml (0.9 .. 0.9999), on `--batch` synthetic Visual-Genome-shaped graphs
per GPU (README: global batch 512 on 4 GPUs = 128 per GPU) with seeded random-init weights.  N > 1 (torchrun, one rank
per GPU): data-parallel replicas, the flat gradient buffer is averaged with NCCL all-reduces over NVLink inside backward.

  value  clean graphs already in HBM;
  e2e    the batch starts in pinned HOST memory each step (H2D inside the timed region) and the step's loss is read back.

`--impl reference`: the UNMODIFIED reference modules staged under oracle/_ref (model, precond, objective, loss) with
torch.optim.Adam on the host cores, batch `--cpu-batch` (ema_pytorch is not in this image: the five moving averages are
lerp_'d by hand the way ema_pytorch does), per-graph rate reported.  bench.py stays the driver's headline (sampling).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench as B  # noqa: E402  (ClockSampler, peaks, emit, model builder)
from diffusesg_b200.utils.synthetic import CONFIGS, synthetic_node_flags, synthetic_state_dict  # noqa: E402

METRIC = "trained scene graphs/sec"
UNIT = "graphs/s"
EMA_COEFS = [0.9, 0.95, 0.99, 0.999, 0.9999]     # config/edm_diffuse_sg/edm_diffuse_sg_regular_visual_genome.yaml:51-56
LR, MAX_NORM = 2.0e-4, 10.0                      # :44, trainer_node_adj.py:174


def clean_batch(cfg, batch, seed):
    """Clean training graphs in the 'bits' encoding: random +-1 bits, U(-1, 1) boxes, masked."""
    g = torch.Generator().manual_seed(seed)
    n, ce, cn = cfg["img"], cfg["c_e"], cfg["c_n"]
    flags = synthetic_node_flags(cfg, batch, seed=seed)
    f = flags.float()
    adj = (torch.randint(0, 2, (batch, ce, n, n), generator=g).float() * 2 - 1) * f[:, None, :, None] * f[:, None, None, :]
    node = torch.randint(0, 2, (batch, n, cn), generator=g).float() * 2 - 1
    node[..., -4:] = torch.rand(batch, n, 4, generator=g) * 2 - 1
    return adj, node * f[:, :, None], flags


def workload(args, cfg, world):
    return {"workload": f"DiffuseSG training step, {cfg['dataset']}-shaped synthetic graphs (N={cfg['img']}, C_e={cfg['c_e']}, "
                        f"C_n={cfg['c_n']}, window {cfg['window']}, depths {cfg['depths']}), EDM loss, self-conditioning coin "
                        f"flip on, clip 10, Adam lr 2e-4, 5 EMAs",
            "batch_per_gpu": args.batch, "global_batch": args.batch * world,
            "parallelism": f"data-parallel x{world}" + (", gradient all-reduce (NCCL) inside backward" if world > 1 else ""),
            "l2": "activations per step (>= 7 GB at batch 128) exceed the 126 MB L2; no flush needed"}


def reference_rate(cfg, cpu_batch, steps, warmup):
    """The unmodified reference's training iteration on all host cores -> dict(rate graphs/s, sec, cores, sample)."""
    from oracle import stage_reference as R
    torch.set_num_threads(os.cpu_count() or 1)
    ref = R.load()
    sd = synthetic_state_dict(cfg, seed=1234, stress=False)
    model = R.build_network(ref, cfg, sd).train()
    model.model.train()
    gen = ref.NodeAdjEDMObjectiveGenerator(precond="edm", sigma_dist="edm", other_params=None, dev="cpu", symmetric_noise=False)
    loss_fn = ref.NodeAdjRainbowLoss(edge_loss_weight=1.0, node_loss_weight=1.0, objective="edm")
    opt = torch.optim.Adam(model.parameters(), lr=LR, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    emas = [[p.detach().clone() for p in model.parameters()] for _ in EMA_COEFS]
    adj, node, flags = clean_batch(cfg, cpu_batch, 1234)
    torch.manual_seed(1234)
    np.random.seed(1234)

    def step(i):
        na, nx, cond, ta, tx, (c_skip, c_out, c_in, c_noise, sigmas, weights) = gen.get_input_output(adj, node, flags)
        opt.zero_grad(set_to_none=True)
        oa, ox = model(adjs=na, nodes=nx, node_flags=flags, sigmas=sigmas)
        la, ln = loss_fn(net_pred_a=oa, net_pred_x=ox, net_target_a=ta, net_target_x=tx, net_cond=cond, adjs_perturbed=na,
                         adjs_gt=adj, x_perturbed=nx, x_gt=node, node_flags=flags, loss_weight=weights, reduction="none")
        (la.mean() + ln.mean()).backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=MAX_NORM, norm_type=2)
        opt.step()
        with torch.no_grad():
            for coef, ema in zip(EMA_COEFS, emas):
                d = min(coef, 1 - 1 / (2 + i))
                for e, p in zip(ema, model.parameters()):
                    e.lerp_(p, 1 - d)

    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step(i)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    sample = (f"the UNMODIFIED reference (oracle/_ref: NodeAdjPrecond(DiffuseSG), objective generator, rainbow loss; "
              f"torch.optim.Adam; hand-written EMA lerps), fp32, dev=cpu, batch {cpu_batch}, {sec:.2f} s per iteration")
    return dict(rate=cpu_batch / sec, sec=sec, cores=torch.get_num_threads(), sample=sample)


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    r = reference_rate(cfg, args.cpu_batch, args.steps, args.warmup)
    config = workload(args, cfg, world)
    config["measured_sample"] = {"batch": args.cpu_batch, "device": "cpu", "world": 1, "sec_per_step": r["sec"]}
    B.emit({"impl": "reference", "metric": METRIC, "value": r["rate"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * r["sec"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": r["rate"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": r["sample"]},
            "e2e": {"value": r["rate"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def run_native(args, cfg, rank, local_rank, world):
    import torch.distributed as dist
    from diffusesg_b200 import native
    from diffusesg_b200.loss.rainbow_loss import NodeAdjRainbowLoss
    from diffusesg_b200.runner.objectives.edm import NodeAdjEDMObjectiveGenerator
    from diffusesg_b200.runner.trainer.trainer_node_adj import train_one_step
    from diffusesg_b200.utils.train_utils import FusedAdam, GraphedTrainStep, NativeDDP, NativeEMA
    if not torch.cuda.is_available():
        raise SystemExit("bench_train.py: no CUDA device - the native path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    native.lib()
    torch.manual_seed(1234 + rank)
    np.random.seed(1234 + rank)
    model = B.build_native_model(cfg, device).train()
    emas = [NativeEMA(model, beta=c) for c in EMA_COEFS]
    opt = FusedAdam(model, lr=LR, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=MAX_NORM)
    opt.attach_emas(emas)
    wrapped = NativeDDP(model) if world > 1 else model
    gen = NodeAdjEDMObjectiveGenerator("edm", "edm", dev=device, symmetric_noise=False)
    loss_fn = NodeAdjRainbowLoss(edge_loss_weight=1.0, node_loss_weight=1.0, objective="edm")
    adj_h, node_h, flags_h = [t.pin_memory() for t in clean_batch(cfg, args.batch, args.data_seed + rank)]
    adj_d, node_d, flags_d = adj_h.to(device), node_h.to(device), flags_h.to(device)
    last = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    graphed = None if args.no_graph else GraphedTrainStep(wrapped, opt, emas, gen, loss_fn, MAX_NORM)

    def step_device():
        if graphed is not None:
            last["loss"] = graphed(adj_d, node_d, flags_d)
        else:
            last["loss"] = train_one_step(wrapped, opt, emas, gen, loss_fn, adj_d, node_d, flags_d, MAX_NORM)

    def step_e2e():
        if graphed is not None:
            la, ln = graphed(adj_h, node_h, flags_h)      # pinned host batch -> static device buffers inside the call
        else:
            a, x, f = adj_h.to(device, non_blocking=True), node_h.to(device, non_blocking=True), flags_h.to(device, non_blocking=True)
            la, ln = train_one_step(wrapped, opt, emas, gen, loss_fn, a, x, f, MAX_NORM)
        last["host_loss"] = float((la.mean() + ln.mean()).item())

    for _ in range(max(3, args.warmup)):
        step_device()
    passes0, launches0 = model.raw_passes, native.launch_count()
    with B.ClockSampler(local_rank) as clocks:
        ms = timed(step_device, args.steps)
    passes = model.raw_passes - passes0
    launches = native.launch_count() - launches0
    with B.ClockSampler(local_rank) as clocks_e2e:
        ms_e2e = timed(step_e2e, args.steps)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    graphs = args.batch * world * args.steps
    value, e2e = graphs / (ms / 1e3), graphs / (ms_e2e / 1e3)
    pk = B.peaks()
    gf = B.GFLOP_PER_PASS.get(args.config)
    # one step = 1 grad forward + backward (2x: dgrad + wgrad) + the no-grad self-conditioning passes actually run
    tf = None if gf is None else (3.0 * args.steps + (passes - args.steps)) * args.batch * gf / 1e3 / (ms / 1e3)
    h2d = int(adj_h.numel() * 4 + node_h.numel() * 4 + flags_h.numel())
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload(args, cfg, world), "gpu_launches": int(launches),
            "schedule": {"cuda_graph_per_iteration": graphed is not None,
                         "note": "gpu_launches counts the library's launch calls; with graphs the ~600 launches of objective + "
                                 "forward + loss + backward replay from one captured graph and only the eager ones are counted"},
            "raw_forward_passes_per_step": passes / args.steps,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "clocks": clocks_e2e.summary()},
            "clocks": clocks.summary(),
            "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                         "frac": None if tf is None else tf / pk["tensor_sustained"], "traffic": None,
                         "note": "whole step per GPU: dense reference flops (forward 13.3 GF/graph, backward 2x, plus the "
                                 "no-grad self-conditioning passes) / step time; peak = " + pk["source"]},
            "loss": float(sum(t.mean() for t in last["loss"]).item())}
    if world == 1 and not args.no_cpu_baseline:
        r = reference_rate(cfg, args.cpu_batch, 2, 1)
        line["cpu_baseline"] = {"value": r["rate"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": r["sample"]}
    B.emit(line)


def main():
    sys.stdout.flush()
    B._RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="vg", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=128, help="graphs per GPU per step")
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per iteration")
    ap.add_argument("--data-seed", type=int, default=1234, help="seed of the synthetic clean batch (+ rank)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
    else:
        run_native(args, cfg, rank, local_rank, world)


if __name__ == "__main__":
    main()
