"""Factory of the sampler and the tolerant checkpoint loader - drop-ins for utils/sampling_utils.py:8 / :34."""
from __future__ import annotations

import logging

from ..runner.mcmc_sampler.edm import NodeAdjEDMSampler


def get_mc_sampler(config):
    flag_clip_samples = config.mcmc.sample_clip.min is not None and config.mcmc.sample_clip.max is not None
    assert config.mcmc.name == "edm"
    mc_sampler = NodeAdjEDMSampler(num_steps=config.mcmc.num_steps, clip_samples=flag_clip_samples,
                                   clip_samples_min=config.mcmc.sample_clip.min,
                                   clip_samples_max=config.mcmc.sample_clip.max,
                                   clip_samples_scope=config.mcmc.sample_clip.scope, dev=config.dev, objective="edm",
                                   self_condition=config.train.self_cond, symmetric_noise=False)
    logging.info("EDM-variant objective. Model: %s. Num of steps: %d", config.mcmc.name, config.mcmc.num_steps)
    logging.info("Self-conditioning: %s", config.train.self_cond)
    return mc_sampler


def load_model(ckp_data, model, weight_keyword):
    """Strict load that tolerates the DP/DDP 'module.' prefix on either side (utils/sampling_utils.py:34-60)."""
    assert weight_keyword in ckp_data
    cur = model.state_dict()
    src = ckp_data[weight_keyword]
    strip = lambda k: k[len("module."):] if k.startswith("module.") else k
    by_bare = {strip(k): v for k, v in src.items()}
    if set(by_bare) != {strip(k) for k in cur}:
        missing = sorted({strip(k) for k in cur} - set(by_bare))[:5]
        extra = sorted(set(by_bare) - {strip(k) for k in cur})[:5]
        raise RuntimeError(f"checkpoint does not match the model: missing {missing}, unexpected {extra}")
    model.load_state_dict({k: by_bare[strip(k)] for k in cur}, strict=True)
    return model
