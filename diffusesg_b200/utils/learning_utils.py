"""Factory of the preconditioned denoiser - the drop-in for utils/learning_utils.py:33 ``get_network``."""
from __future__ import annotations

import logging
import os

import torch

from ..model.diffusesg.diffusesg import DiffuseSG
from ..model.precond.precond import NodeAdjPrecond
from .sg_utils import get_node_adj_model_input_output_channels


def get_network(config, dist_helper=None):
    """Same config keys as the reference: dataset.max_node_num, dataset.name,
    model.{name, feature_dims, depths, window_size, patch_size}, train.self_cond, mcmc.{name, precond},
    flag_sg, dev, logdir, train.resume."""
    model_config = config.model
    if model_config.name not in ["diffuse_sg"]:
        raise ValueError(f"Unknown model name {model_config.name}")
    if config.mcmc.name != "edm" or not config.flag_sg:
        raise NotImplementedError("only the EDM scene-graph path (mcmc.name == 'edm', flag_sg) is built")
    feature_nums = model_config.feature_dims if "feature_dims" in model_config else [0]
    in_chans, out_chans_adj, out_chans_node = get_node_adj_model_input_output_channels(config)
    denoising_model = DiffuseSG(img_size=config.dataset.max_node_num, in_chans=in_chans,
                                patch_size=model_config.patch_size, embed_dim=feature_nums[-1],
                                depths=model_config.depths, num_heads=[3, 6, 12, 24],
                                window_size=model_config.window_size, mlp_ratio=4., drop_rate=0., attn_drop_rate=0.,
                                drop_path_rate=0.0, self_condition=config.train.self_cond,
                                symmetric_noise=not config.flag_sg, out_chans_adj=out_chans_adj,
                                out_chans_node=out_chans_node).to(config.dev)
    denoising_model = NodeAdjPrecond(precond=config.mcmc.precond, model=denoising_model,
                                     self_condition=config.train.self_cond, symmetric_noise=not config.flag_sg)
    denoising_model.plot_save_dir = os.path.join(getattr(config, "logdir", "."), "training_plot")
    n_params = sum(p.numel() for p in denoising_model.parameters())
    logging.info(f"Parameters Count: {n_params:,}")
    resume = getattr(config.train, "resume", None)
    if resume is not None:
        from .sampling_utils import load_model
        denoising_model = load_model(torch.load(resume, map_location="cpu"), denoising_model, "model")
    if dist_helper is not None:
        denoising_model = dist_helper.dist_adapt_model(denoising_model)
    return denoising_model


def get_optimizer(model, config, dist_helper=None):
    """utils/learning_utils.py:126-145: Adam(lr_init, betas (0.9, 0.999), eps 1e-8, weight_decay) + ExponentialLR.  The
    reference shards the optimiser state with ZeroRedundancyOptimizer under DDP; here the 36 M-parameter state is one
    flat buffer per GPU and the step is one fused launch (utils/train_utils.py)."""
    from .train_utils import FusedAdam
    optimizer = FusedAdam(model, lr=config.train.lr_init, betas=(0.9, 0.999), eps=1e-8,
                          weight_decay=config.train.weight_decay, max_grad_norm=10.0)
    scheduler = torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=config.train.lr_dacey)
    return optimizer, scheduler


def get_ema_helper(config, model, optimizer=None):
    """utils/learning_utils.py:148-166: one moving average per coefficient (sorted), or None.  Pass the FusedAdam to have
    every average updated inside the optimiser launch."""
    from .train_utils import FusedAdam, NativeEMA
    ema_coef = config.train.ema_coef
    flag_ema = isinstance(ema_coef, list) or (isinstance(ema_coef, float) and ema_coef < 1)
    if not flag_ema:
        logging.info("Exponential moving average is OFF.")
        return None
    coefs = [ema_coef] if isinstance(ema_coef, float) else list(ema_coef)
    inner = model.module if hasattr(model, "module") else model
    helpers = [NativeEMA(inner, beta=c) for c in sorted(coefs)]
    if isinstance(optimizer, FusedAdam):
        optimizer.attach_emas(helpers)
    logging.info("Exponential moving average is ON. Coefficient: {}".format(coefs))
    return helpers
