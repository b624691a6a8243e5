"""Optimiser, moving averages and data-parallel wrapper of the training step, over the flat parameter
buffer of model/diffusesg/train_graph.py.

Drop-ins for utils/learning_utils.py:126-166 (``get_optimizer``: Adam(lr_init, betas (0.9, 0.999), eps 1e-8,
weight_decay) + ExponentialLR; ``get_ema_helper``: one ``ema_pytorch.EMA(beta=coef, update_every=1,
update_after_step=0, inv_gamma=1, power=1)`` per coefficient) and utils/dist_training.py:62-69 (DDP wrap).  One fused
launch (dsg_tr_adam_ema) applies gradient clipping (trainer_node_adj.py:174), Adam and every moving average; the
gradient all-reduce is NCCL over NVLink on two contiguous ranges of the flat gradient buffer, overlapped with backward.
"""
from __future__ import annotations

import copy
import ctypes as C
from typing import List, Optional

import torch
import torch.nn as nn

from .. import native
from ..model.diffusesg.diffusesg import DiffuseSG
from ..model.diffusesg.train_graph import train_state


def find_denoiser(model: nn.Module) -> DiffuseSG:
    for m in model.modules():
        if isinstance(m, DiffuseSG):
            return m
    raise ValueError("no DiffuseSG module inside the model")


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (L2 weight decay added to the gradient, bias-corrected moments) on the flat parameter
    buffer, optional clip_grad_norm_ (`max_grad_norm`) and the attached moving averages, in one kernel launch."""

    def __init__(self, model: nn.Module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=None):
        self.denoiser = find_denoiser(model)
        dev = next(self.denoiser.parameters()).device
        self.ts = train_state(self.denoiser, dev)
        params = [p for p in model.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        n_flat = sum(1 for _ in self.denoiser.parameters())
        if len(params) != n_flat:
            raise NotImplementedError("FusedAdam: the model has trainable parameters outside the DiffuseSG denoiser")
        self.max_grad_norm = max_grad_norm
        with native.device_guard(dev):
            self.m = torch.zeros_like(self.ts.flat)
            self.v = torch.zeros_like(self.ts.flat)
            self.gsumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.steps = 0
        self.emas: List["NativeEMA"] = []

    # checkpoint / resume (the reference's trainer saves optimizer.state_dict() next to the model): the two moment buffers
    # are stored per parameter under the parameter's name, so a checkpoint does not depend on the flat layout
    def state_dict(self):
        ts = self.ts
        per = {k: dict(exp_avg=self.m[ts.offs[k]: ts.offs[k] + ts.params[k].numel()].view(ts.shapes[k]).clone(),
                       exp_avg_sq=self.v[ts.offs[k]: ts.offs[k] + ts.params[k].numel()].view(ts.shapes[k]).clone())
               for k in ts.order}
        groups = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        return {"fused_adam": 1, "step": self.steps, "state": per, "param_groups": groups}

    def load_state_dict(self, sd):
        if "fused_adam" not in sd:
            raise ValueError("FusedAdam.load_state_dict: not a FusedAdam checkpoint")
        ts = self.ts
        for k in ts.order:
            n = ts.params[k].numel()
            self.m[ts.offs[k]: ts.offs[k] + n].copy_(sd["state"][k]["exp_avg"].reshape(-1))
            self.v[ts.offs[k]: ts.offs[k] + n].copy_(sd["state"][k]["exp_avg_sq"].reshape(-1))
        self.steps = int(sd["step"])
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            g.update(saved)

    def attach_emas(self, emas):
        self.emas = list(emas or [])
        for e in self.emas:
            e.fused_into = self

    @torch.no_grad()
    def step(self, closure=None):
        ts = self.ts
        if not ts.attached():
            raise native.NativeError("FusedAdam: the model's parameters were moved off the flat buffer (.to() after "
                                     "construction?); rebuild the optimiser")
        if ts.attach_grads():
            return None   # no backward ran since zero_grad(set_to_none=True): nothing to apply
        g = self.param_groups[0]
        self.steps += 1
        st = native.stream_ptr(ts.dev)
        lib = native.lib()
        with native.device_guard(ts.dev):
            gs = None
            if self.max_grad_norm is not None:
                native.check(lib.dsg_tr_sumsq(ts.grad.data_ptr(), ts.numel, self.gsumsq.data_ptr(), st), "dsg_tr_sumsq")
                gs = self.gsumsq.data_ptr()
            n = len(self.emas)
            ptrs = (C.c_void_p * 8)(*[e.flat.data_ptr() for e in self.emas])
            decays = (C.c_float * 8)(*[e.next_decay() for e in self.emas])
            native.check(lib.dsg_tr_adam_ema(ts.flat.data_ptr(), ts.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                             ts.numel, gs, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                                             float(g["eps"]), float(g["weight_decay"]), self.steps,
                                             float(self.max_grad_norm or 0.0), n, ptrs, decays, st), "dsg_tr_adam_ema")
        for e in self.emas:
            e.fused_done()
        self.denoiser.invalidate_native()
        return None

    def grad_norm(self) -> torch.Tensor:
        """sqrt of the sum of squares the last step clipped with (device scalar)."""
        return self.gsumsq.sqrt()


class NativeEMA:
    """ema_pytorch.EMA(model, beta, update_every=1, update_after_step=0, inv_gamma=1, power=1) on a flat buffer:
    ``.ema_model`` is a deep copy of the online model whose denoiser parameters are views of ``.flat``; ``update()``
    follows ema_pytorch's schedule (copy on the first two calls, then decay = min(beta, 1 - 1 / (1 + epoch)))."""

    def __init__(self, model: nn.Module, beta: float):
        self.online = model
        self.beta = float(beta)
        self.online_ts = train_state(find_denoiser(model), next(model.parameters()).device)
        self.ema_model = copy.deepcopy(model)
        self.ema_model.requires_grad_(False)
        den = find_denoiser(self.ema_model)
        self.ts = train_state(den, self.online_ts.dev)
        self.flat = self.ts.flat
        self.denoiser = den
        self.step = 0
        self.initted = False
        self.fused_into: Optional[FusedAdam] = None
        self._pending = 0

    def next_decay(self) -> float:
        """Decay ema_pytorch would use for the update that follows the current optimiser step."""
        step = self.step
        if step <= 0 or not self.initted:
            return 0.0                      # copy_params_from_model_to_ema
        epoch = max(step + 1 - 0 - 1, 0)    # get_current_decay reads self.step AFTER the increment
        if epoch <= 0:
            return 0.0
        return min(max(1.0 - 1.0 / (1.0 + epoch), 0.0), self.beta)

    def _advance(self):
        if self.step > 0:
            self.initted = True
        self.step += 1
        self.denoiser.invalidate_native()

    def fused_done(self):
        self._advance()
        self._pending += 1

    def update(self):
        """trainer_node_adj.py:178.  A no-op when the optimiser's fused launch already moved this average."""
        if self._pending > 0:
            self._pending -= 1
            return
        d = self.next_decay()
        lib = native.lib()
        ptrs = (C.c_void_p * 8)(self.flat.data_ptr())
        decays = (C.c_float * 8)(d)
        with native.device_guard(self.ts.dev):
            native.check(lib.dsg_tr_adam_ema(self.online_ts.flat.data_ptr(), None, None, None, self.online_ts.numel, None,
                                             0.0, 0.9, 0.999, 1e-8, 0.0, 1, 0.0, 1, ptrs, decays,
                                             native.stream_ptr(self.ts.dev)), "dsg_tr_adam_ema")
        self._advance()


class NativeDDP(nn.Module):
    """Data-parallel wrapper of the training step (utils/dist_training.py:62-69 wraps with torch DDP): parameters are
    broadcast from rank 0 at construction; the tape's backward averages the flat gradient buffer over the group with
    NCCL all-reduces (the read-out range while the U-Net backward still runs, the rest at the end)."""

    def __init__(self, module: nn.Module, process_group=None):
        super().__init__()
        import torch.distributed as dist
        self.module = module
        den = find_denoiser(module)
        ts = train_state(den, next(den.parameters()).device)
        dist.broadcast(ts.flat, src=0, group=process_group)
        ts.ddp_group = process_group if process_group is not None else dist.group.WORLD
        den.invalidate_native()

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


class GraphedTrainStep:
    """The training iteration of runner/trainer/trainer_node_adj.py:95-178 with everything between the clean batch and the
    gradients - objective generator, preconditioned denoiser (both self-conditioning outcomes), loss, backward: ~600 kernel
    launches - replayed as ONE CUDA graph; gradient all-reduce (DDP), optimiser and moving averages follow eagerly (three
    launches).  The self-conditioning coin is drawn here from the global numpy stream, exactly where NodeAdjPrecond.forward
    would draw it, and selects one of two captured graphs; the torch CUDA generator advances per replay as in eager mode.
    Shapes are fixed at the first call."""

    def __init__(self, model, optimizer, ema_helper, train_obj_gen, loss_func, max_grad_norm: float = 10.0, warmup: int = 2):
        import numpy as np
        self.np = np
        self.model, self.opt, self.emas = model, optimizer, ema_helper
        self.gen, self.loss_func, self.max_grad_norm = train_obj_gen, loss_func, max_grad_norm
        self.precond = model.module if hasattr(model, "module") else model
        self.den = find_denoiser(model)
        self.warmup = warmup
        self.graphs = {}
        self.static = None
        self.pool = None

    def _body(self):
        adj, node, flags = self.static
        na, nx, cond, ta, tx, (c_skip, c_out, c_in, c_noise, sigmas, weights) = self.gen.get_input_output(adj, node, flags)
        for p in self.den.parameters():
            p.grad = None                      # zero_grad(set_to_none=True): the backward zeroes the flat buffer itself
        oa, ox = self.precond(adjs=na, nodes=nx, node_flags=flags, sigmas=sigmas)
        la, ln = self.loss_func(net_pred_a=oa, net_pred_x=ox, net_target_a=ta, net_target_x=tx, net_cond=cond,
                                adjs_perturbed=na, adjs_gt=adj, x_perturbed=nx, x_gt=node, node_flags=flags,
                                loss_weight=weights, reduction="none")
        (la.mean() + ln.mean()).backward()
        return la.detach(), ln.detach()

    def _capture(self, coin: bool):
        ts = train_state(self.den, self.static[0].device)
        group, ts.ddp_group = ts.ddp_group, None      # the all-reduce runs after the replay, outside the graph
        self.precond.__dict__["_forced_coin"] = coin
        dev = self.static[0].device
        rng = torch.cuda.get_rng_state(dev)           # warm-up and capture must not consume the step's noise draws
        try:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(self.warmup):
                    self._body()
            cur.wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self.pool):
                out = self._body()
            if self.pool is None:
                self.pool = g.pool()
        finally:
            self.precond.__dict__["_forced_coin"] = None
            ts.ddp_group = group
            torch.cuda.set_rng_state(rng, dev)
        self.graphs[coin] = (g, out)

    def __call__(self, adjs_gt, nodes_gt, node_flags):
        dev = self.gen.dev
        if self.static is None:
            self.static = (torch.empty_like(adjs_gt, device=dev), torch.empty_like(nodes_gt, device=dev),
                           torch.empty_like(node_flags, device=dev))
        for dst, src in zip(self.static, (adjs_gt, nodes_gt, node_flags)):
            if dst.shape != src.shape or dst.dtype != src.dtype:
                raise ValueError(f"GraphedTrainStep: batch of shape {tuple(src.shape)} / {src.dtype}, the graphs were captured "
                                 f"for {tuple(dst.shape)} / {dst.dtype} (fixed at the first call; use train_one_step for "
                                 "ragged last batches)")
            dst.copy_(src, non_blocking=True)
        coin = bool(self.precond.self_condition and self.np.random.rand() < 0.5)
        passes0 = self.precond.raw_passes
        if coin not in self.graphs:
            self._capture(coin)                # warm-up + capture only; the replay below computes this batch's gradients
        g, (la, ln) = self.graphs[coin]
        g.replay()
        self.precond.raw_passes = passes0 + (2 if coin else 1)
        ts = train_state(self.den, self.static[0].device)
        if ts.ddp_group is not None:
            import torch.distributed as dist
            dist.all_reduce(ts.grad, op=dist.ReduceOp.AVG, group=ts.ddp_group)
        if isinstance(self.opt, FusedAdam):
            self.opt.max_grad_norm = self.max_grad_norm
        else:
            nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self.max_grad_norm, norm_type=2)
        self.opt.step()
        if self.emas is not None:
            for ema in self.emas:
                ema.update()
        return la, ln
