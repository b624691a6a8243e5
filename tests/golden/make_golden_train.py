"""Golden vectors of ONE training iteration's forward + backward, produced by executing the UNMODIFIED reference
(build container only: needs /root/reference).

    python tests/golden/make_golden_train.py

The reference's own modules - DiffuseSG, NodeAdjPrecond, NodeAdjRainbowLoss - run the step of
runner/trainer/trainer_node_adj.py:104-173 (model call, loss with reduction='none' and loss weights, mean + mean,
backward) on CPU in fp32 with the seeded synthetic weights / inputs of diffusesg_b200.utils.synthetic; committed are the
two per-sample losses, the L2 norm and the sum of EVERY parameter's gradient, and a handful of complete gradient tensors.
tests/test_oracle_golden.py holds the oracle's autograd to them (CPU), tests/test_gpu_train.py the native backward (B200).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

FULL = ["patch_embed.proj.weight", "patch_embed.norm.weight", "patch_embed.affine.bias", "map_layer0.weight",
        "down_layers.0.blocks.0.norm1.weight", "down_layers.0.blocks.0.attn.relative_position_bias_table",
        "down_layers.0.blocks.0.attn.qkv.bias", "down_layers.0.blocks.0.attn.proj.weight", "down_layers.0.blocks.0.mlp.fc2.bias",
        "down_layers.0.downsample.norm.bias", "down_layers.1.blocks.0.affine.bias", "up_layers.1.upsample.post_norm.weight",
        "up_layers.1.upsample.post_linear.weight", "read_out.0.weight", "read_out.2.bias", "norm.bias",
        "readout_adj_mlp.fc2.weight", "readout_node_mlp.fc1.weight", "readout_node_mlp.fc2.bias"]


def case_inputs(cfg, batch):
    """Shared with the tests: noisy inputs, clean targets, sigmas, loss weights, self-conditioning inputs."""
    from diffusesg_b200.utils.synthetic import synthetic_inputs
    adj, node, flags, _, sc_adj, sc_node = synthetic_inputs(cfg, batch, seed=11)
    sigmas = torch.tensor([0.2, 1.5, 4.0, 0.7, 0.05, 9.0, 0.9, 2.2])[:batch].contiguous()
    f = flags.float()
    tgt_adj = adj.sign() * f[:, None, :, None] * f[:, None, None, :]
    tgt_node = node.clamp(-1, 1) * f[:, :, None]
    weights = (sigmas ** 2 + 0.25) / (sigmas * 0.5) ** 2
    return adj, node, flags, sigmas, sc_adj, sc_node, tgt_adj, tgt_node, weights


def main():
    sys.path.insert(0, HERE)
    from make_golden import build_ref, import_reference
    from diffusesg_b200.utils.synthetic import CONFIGS, synthetic_state_dict
    torch.set_num_threads(8)
    DiffuseSG, NodeAdjPrecond, _ = import_reference()
    sys.path.insert(0, "/root/reference/DiffuseSG")
    from loss.rainbow_loss import NodeAdjRainbowLoss
    sys.path.remove("/root/reference/DiffuseSG")
    for name, batch in (("tiny", 4), ("vg", 2)):
        cfg = CONFIGS[name]
        net = build_ref(DiffuseSG, cfg)
        net.load_state_dict(synthetic_state_dict(cfg, seed=1234, stress=True), strict=True)
        # self_condition=False on the wrapper: no coin flip, the self-conditioning inputs below still reach the network
        model = NodeAdjPrecond(precond="edm", model=net, self_condition=False, symmetric_noise=False).train()
        loss_fn = NodeAdjRainbowLoss(edge_loss_weight=1.0, node_loss_weight=1.0, objective="edm")
        adj, node, flags, sigmas, sc_adj, sc_node, tgt_adj, tgt_node, weights = case_inputs(cfg, batch)
        oa, ox = model(adjs=adj, nodes=node, node_flags=flags, sigmas=sigmas, self_cond_adjs=sc_adj, self_cond_nodes=sc_node)
        la, ln = loss_fn(net_pred_a=oa, net_pred_x=ox, net_target_a=tgt_adj, net_target_x=tgt_node, net_cond=sigmas,
                         adjs_perturbed=adj, adjs_gt=tgt_adj, x_perturbed=node, x_gt=tgt_node, node_flags=flags,
                         loss_weight=weights, reduction="none")
        (la.mean() + ln.mean()).backward()
        out = {"loss_adj": la.detach().numpy(), "loss_node": ln.detach().numpy()}
        keys, norms, sums = [], [], []
        for k, p in net.named_parameters():
            keys.append(k)
            norms.append(float(p.grad.double().norm()))
            sums.append(float(p.grad.double().sum()))
            if k in FULL and (name == "tiny" or p.numel() <= 20000):
                out["grad::" + k] = p.grad.numpy().copy()
        out["keys"] = np.array(keys)
        out["norms"] = np.array(norms)
        out["sums"] = np.array(sums)
        np.savez_compressed(os.path.join(HERE, f"train_grads_{name}.npz"), **out)
        print(name, "loss", float(la.mean() + ln.mean()), "params", len(keys), "full tensors", sum(k.startswith("grad::") for k in out))


if __name__ == "__main__":
    main()
