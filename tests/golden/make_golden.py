"""Generate golden vectors by executing the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py            # needs /root/reference

The reference modules are imported from where they lie (nothing is copied into
the repo); the two packages they import at module level but which are absent
from this image (timm: DropPath/to_2tuple/trunc_normal_) are
replaced by three-line stand-ins in ``sys.modules``.  Weights and inputs come
from ``diffusesg_b200.utils.synthetic`` (seeded, regenerable anywhere), so only
OUTPUTS are committed, as small ``.npz`` / ``.json`` files next to this script.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/DiffuseSG"
sys.path.insert(0, ROOT)


def import_reference():
    layers = types.ModuleType("timm.models.layers")

    class DropPath(torch.nn.Identity):
        def __init__(self, p=0.0):
            super().__init__()

    layers.DropPath = DropPath
    layers.to_2tuple = lambda v: v if isinstance(v, tuple) else (v, v)
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    sys.modules["timm"] = types.ModuleType("timm")
    sys.modules["timm.models"] = types.ModuleType("timm.models")
    sys.modules["timm.models.layers"] = layers
    sys.path.insert(0, REF)
    from model.diffusesg.diffusesg import DiffuseSG
    from model.precond.precond import NodeAdjPrecond
    from runner.mcmc_sampler.edm import NodeAdjEDMSampler
    sys.path.remove(REF)
    return DiffuseSG, NodeAdjPrecond, NodeAdjEDMSampler


def build_ref(DiffuseSG, cfg):
    from diffusesg_b200.utils.synthetic import in_chans
    return DiffuseSG(img_size=cfg["img"], in_chans=in_chans(cfg), patch_size=1, embed_dim=cfg["embed"],
                     depths=cfg["depths"], num_heads=[3, 6, 12, 24], window_size=cfg["window"], mlp_ratio=4.,
                     drop_rate=0., attn_drop_rate=0., drop_path_rate=0.0, self_condition=cfg["self_cond"],
                     symmetric_noise=False, out_chans_adj=cfg["c_e"], out_chans_node=cfg["c_n"])


def main():
    torch.set_num_threads(8)
    from diffusesg_b200.utils.synthetic import CONFIGS, synthetic_inputs, synthetic_state_dict
    DiffuseSG, NodeAdjPrecond, NodeAdjEDMSampler = import_reference()

    # ---- (1) state_dict layout and the reference's own seeded init ------------------------------
    for name in ("vg", "coco", "tiny", "n64w16"):
        cfg = CONFIGS[name]
        torch.manual_seed(1234)
        net = build_ref(DiffuseSG, cfg)
        sd = net.state_dict()
        layout = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()]
        sums = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}
        with open(os.path.join(HERE, f"layout_{name}.json"), "w") as f:
            json.dump(dict(layout=layout, n_params=sum(p.numel() for p in net.parameters()),
                           init_seed=1234, init_sums=sums), f)
        print(name, len(layout), "keys")

    # ---- (2) raw network forward ----------------------------------------------------------------
    for name, batch in (("tiny", 3), ("vg", 2), ("coco", 2)):
        cfg = CONFIGS[name]
        net = build_ref(DiffuseSG, cfg).eval()
        net.load_state_dict(synthetic_state_dict(cfg, seed=1234, stress=True), strict=True)
        adj, node, flags, sigmas, sc_adj, sc_node = synthetic_inputs(cfg, batch, seed=7)
        sig = sigmas * torch.linspace(0.5, 2.0, batch)
        labels = sig.log() / 4
        out = {}
        with torch.no_grad():
            a, n = net(adj.clone(), node.clone(), flags, labels, sc_adj.clone(), sc_node.clone())
            out["adj_sc"], out["node_sc"] = a.numpy(), n.numpy()
            a, n = net(adj.clone(), node.clone(), flags, labels, None, None)
            out["adj_nosc"], out["node_nosc"] = a.numpy(), n.numpy()
        out["labels"] = labels.numpy()
        np.savez_compressed(os.path.join(HERE, f"forward_{name}.npz"), **out)
        print("forward", name, {k: v.shape for k, v in out.items()})

    # ---- (3) preconditioned call with the coin flip ---------------------------------------------
    cfg = CONFIGS["tiny"]
    net = build_ref(DiffuseSG, cfg).eval()
    net.load_state_dict(synthetic_state_dict(cfg, seed=1234, stress=True), strict=True)
    model = NodeAdjPrecond(precond="edm", model=net, self_condition=True, symmetric_noise=False).eval()
    adj, node, flags, sigmas, sc_adj, sc_node = synthetic_inputs(cfg, 3, seed=7)
    out = {}
    np.random.seed(5)
    coins = np.random.rand(4)
    np.random.seed(5)
    sa = sn = None
    with torch.no_grad():
        for k, s in enumerate((40.0, 3.0, 0.4, 0.01)):
            sg = torch.full((3,), s)
            sa, sn = model(adj * s, node * s, flags, sg, sa, sn)
            out[f"adj_{k}"], out[f"node_{k}"] = sa.numpy().copy(), sn.numpy().copy()
    out["coins"] = coins
    np.savez_compressed(os.path.join(HERE, "precond_tiny.npz"), **out)
    print("precond coins", coins)

    # ---- (4) sampler trajectories ---------------------------------------------------------------
    steps = 8
    sampler = NodeAdjEDMSampler(num_steps=steps, clip_samples=True, clip_samples_min=-1.0, clip_samples_max=1.0,
                                clip_samples_scope="x_0", dev="cpu", objective="edm", self_condition=True,
                                symmetric_noise=False)
    flags2 = flags[:2]
    torch.manual_seed(11)
    np.random.seed(11)
    a, n, a_ls, n_ls = sampler.sample(model=model, node_flags=flags2, flag_interim_adjs=True,
                                      max_num_interim_adjs=4, num_node_chan=cfg["c_n"], num_edge_chan=cfg["c_e"])
    out = dict(adjs=a.numpy(), nodes=n.numpy(), adjs_ls=a_ls.numpy(), nodes_ls=n_ls.numpy(),
               t_steps=torch.cat([sampler.sigma_steps, torch.zeros(1, dtype=torch.float64)]).to(torch.float32).numpy())
    # known-answer mode of the reference (edm.py:372-377): the sampler must return the ground truth
    gt_a, gt_n = adj[:2].sign() * (flags2[:, None, :, None] & flags2[:, None, None, :]), node[:2].clamp(-1, 1)
    torch.manual_seed(12)
    ka, kn = sampler.sample(model=model, node_flags=flags2, sanity_check_gt_adjs=gt_a, sanity_check_gt_nodes=gt_n,
                            num_node_chan=cfg["c_n"], num_edge_chan=cfg["c_e"])
    out.update(kat_adjs=ka.numpy(), kat_nodes=kn.numpy(), kat_gt_adjs=gt_a.numpy(), kat_gt_nodes=gt_n.numpy())
    # the 256-step schedule the shipped configs use
    full = NodeAdjEDMSampler(num_steps=256, clip_samples=True, clip_samples_min=-1.0, clip_samples_max=1.0,
                             clip_samples_scope="x_0", dev="cpu", objective="edm", self_condition=True,
                             symmetric_noise=False)
    out["sigma_steps_256"] = full.sigma_steps.numpy()
    np.savez_compressed(os.path.join(HERE, "sampler_tiny.npz"), **out)
    print("sampler", a.shape, n.shape, a_ls.shape, n_ls.shape, "KAT max err",
          float((ka - gt_a).abs().max()), float((kn - gt_n).abs().max()))


def make_train_golden():
    """(5) training objective + loss of the unmodified reference, tiny geometry, batch 4."""
    torch.set_num_threads(8)
    from diffusesg_b200.utils.synthetic import CONFIGS, synthetic_inputs
    import_reference()
    sys.path.insert(0, REF)
    from runner.objectives.edm import NodeAdjEDMObjectiveGenerator
    from loss.rainbow_loss import NodeAdjRainbowLoss
    sys.path.remove(REF)
    cfg = CONFIGS["tiny"]
    adj, node, flags, _, _, _ = synthetic_inputs(cfg, 4, seed=7)
    pair = flags[:, None, :, None] & flags[:, None, None, :]
    clean_a = adj.sign() * pair                                   # +-1 bits, masked
    clean_x = node.clamp(-1, 1) * flags[:, :, None]
    gen = NodeAdjEDMObjectiveGenerator("edm", "edm", other_params=None, dev="cpu", symmetric_noise=False)
    torch.manual_seed(31)
    in_a, in_x, cond, tgt_a, tgt_x, (c_skip, c_out, c_in, c_noise, sigmas, weights) = gen.get_input_output(
        clean_a, clean_x, flags)
    # replay the draws the call above consumed: randn(B), randn_like(adj), randn_like(node)
    torch.manual_seed(31)
    rnd, eps_a, eps_x = torch.randn(4), torch.randn_like(clean_a), torch.randn_like(clean_x)
    # a stand-in prediction: target + a smooth perturbation (no network involved)
    torch.manual_seed(32)
    pred_a = _maskf(tgt_a + 0.3 * torch.randn_like(tgt_a), pair)
    pred_x = (tgt_x + 0.2 * torch.randn_like(tgt_x)) * flags[:, :, None]
    loss = NodeAdjRainbowLoss(edge_loss_weight=1.0, node_loss_weight=0.5, objective="edm")
    out = dict(clean_a=clean_a.numpy(), clean_x=clean_x.numpy(), rnd=rnd.numpy(), eps_a=eps_a.numpy(),
               eps_x=eps_x.numpy(), in_a=in_a.numpy(), in_x=in_x.numpy(), sigmas=sigmas.numpy(),
               weights=weights.numpy(), c_skip=c_skip.numpy(), c_out=c_out.numpy(), c_in=c_in.numpy(),
               c_noise=c_noise.numpy(), pred_a=pred_a.numpy(), pred_x=pred_x.numpy())
    for red in ("none", "mean"):
        la, ln = loss(pred_a, pred_x, tgt_a, tgt_x, cond, in_a, clean_a, in_x, clean_x, flags, loss_weight=weights,
                      reduction=red)
        out[f"loss_adj_{red}"], out[f"loss_node_{red}"] = la.numpy(), ln.numpy()
    np.savez_compressed(os.path.join(HERE, "train_objective.npz"), **out)
    print("train objective", {k: v.shape for k, v in out.items()})


def _maskf(t, m):
    return torch.where(m, t, torch.zeros_like(t))


if __name__ == "__main__":
    if "--only-train" in sys.argv:
        make_train_golden()
    else:
        main()
        make_train_golden()
