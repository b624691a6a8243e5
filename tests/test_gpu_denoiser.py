"""End-to-end parity of the native denoiser / preconditioner / sampler against the oracle and the golden
vectors of the unmodified reference (run on the B200 with -m gpu).

Stated tolerances (bf16 tensor-core operands, fp32 accumulation / LayerNorm / softmax / residual stream):
  raw network output F:   rel-L2 <= 2e-2 per pass   (a bf16-autocast run of the reference itself is at 1.3e-2)
  D = c_skip x + c_out F: |err| <= 2e-2 * c_out * rms(F) + fp32 slack
"""
import os

import numpy as np
import pytest
import torch

from diffusesg_b200 import native
from diffusesg_b200.model.diffusesg.diffusesg import DiffuseSG
from diffusesg_b200.model.precond.precond import NodeAdjPrecond
from diffusesg_b200.runner.mcmc_sampler.edm import NodeAdjEDMSampler
from diffusesg_b200.utils.synthetic import CONFIGS, in_chans, synthetic_inputs, synthetic_state_dict
from oracle import denoiser_oracle as O
from oracle import edm_oracle as E

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
F_TOL = 2e-2


def build(cfg, stress=True):
    m = DiffuseSG(img_size=cfg["img"], in_chans=in_chans(cfg), patch_size=1, embed_dim=cfg["embed"],
                  depths=cfg["depths"], num_heads=[3, 6, 12, 24], window_size=cfg["window"], mlp_ratio=4.,
                  drop_rate=0., attn_drop_rate=0., drop_path_rate=0.0, self_condition=cfg["self_cond"],
                  symmetric_noise=False, out_chans_adj=cfg["c_e"], out_chans_node=cfg["c_n"])
    sd = synthetic_state_dict(cfg, seed=1234, stress=stress)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


def oracle_net(cfg, sd, capture=None):
    def f(adj, node, flags, labels, sa, sn):
        return O.denoiser_forward(sd, img=cfg["img"], embed=cfg["embed"], depths=cfg["depths"], heads=cfg["heads"],
                                  window=cfg["window"], self_condition=cfg["self_cond"], adj=adj, node=node,
                                  flags=flags, noise_labels=labels, sc_adj=sa, sc_node=sn, capture=capture)
    return f


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30))


def stage_names(cfg):
    nl, d = len(cfg["depths"]), cfg["depths"]
    names = ["patch_embed"]
    for s in range(nl):
        names += [f"down_layers.{s}.blocks.{j}" for j in range(d[s])]
        if s < nl - 1:
            names.append(f"down_layers.{s}.downsample")
    for u in range(nl):
        if u > 0:
            names.append(f"up_layers.{u}.upsample")
        names += [f"up_layers.{u}.blocks.{j}" for j in range(d[nl - 1 - u])]
    return names


@pytest.mark.parametrize("name,batch", [("tiny", 3), ("vg", 2), ("coco", 2), ("n64w16", 2)])
def test_forward_stage_by_stage(name, batch):
    """Localises a parity break: leave the native schedule after every stage and compare the residual stream
    with the oracle's capture of the same stage."""
    cfg = CONFIGS[name]
    model, sd = build(cfg)
    adj, node, flags, sigmas, sc_adj, sc_node = synthetic_inputs(cfg, batch, seed=7)
    labels = (sigmas * torch.linspace(0.5, 2.0, batch)).log() / 4
    cap = {}
    with torch.no_grad():
        oracle_net(cfg, sd, cap)(adj, node, flags, labels, sc_adj, sc_node)
    lib = native.lib()
    report = []
    try:
        for k, stage in enumerate(stage_names(cfg)):
            lib.dsg_debug_set_stop_after(k)
            with torch.no_grad():
                model(adj.to(DEV), node.to(DEV), flags.to(DEV), labels.to(DEV), sc_adj.to(DEV), sc_node.to(DEV))
            torch.cuda.synchronize()
            want = cap[stage]
            buf = "X"
            if stage.endswith("downsample"):
                buf = "skip" + stage.split(".")[1]
            got = model._nat.debug_buffer(buf, torch.float32)[:want.numel()].view(want.shape).cpu()
            report.append((stage, rel(got, want)))
    finally:
        lib.dsg_debug_set_stop_after(-1)
    msg = "\n".join(f"{s:40s} {e:.3e}" for s, e in report)
    print(msg)
    assert all(e < F_TOL for _, e in report), "\n" + msg


@pytest.mark.parametrize("name,batch", [("tiny", 3), ("vg", 2), ("coco", 2)])
def test_forward_matches_golden(name, batch, golden_dir):
    cfg = CONFIGS[name]
    g = np.load(os.path.join(golden_dir, f"forward_{name}.npz"))
    model, _ = build(cfg)
    adj, node, flags, sigmas, sc_adj, sc_node = [t.to(DEV) for t in synthetic_inputs(cfg, batch, seed=7)]
    labels = torch.from_numpy(g["labels"]).to(DEV)
    with torch.no_grad():
        a, n = model(adj, node, flags, labels, sc_adj, sc_node)
        a0, n0 = model(adj, node, flags, labels, None, None)
    errs = {key: rel(got, torch.from_numpy(g[key])) for got, key in
            ((a, "adj_sc"), (n, "node_sc"), (a0, "adj_nosc"), (n0, "node_nosc"))}
    print(name, errs)
    for got in (a, n, a0, n0):
        assert torch.isfinite(got).all()
    assert all(e < F_TOL for e in errs.values()), errs
    # masking is exact
    pair = (flags[:, None, :, None] & flags[:, None, None, :]).expand_as(a)
    assert float(a[~pair].abs().sum()) == 0.0 and float(n[~flags].abs().sum()) == 0.0


@pytest.mark.parametrize("name,batch", [("tiny", 5), ("vg", 3)])
def test_fused_mlp_matches_unfused_schedule(name, batch):
    """The fused LN + fc1 + GELU + fc2 kernel against the LayerNorm kernel + two GEMMs it replaces (same bf16
    roundings of the LN output and of the hidden activation; only the fp32 accumulation order differs)."""
    cfg = CONFIGS[name]
    inputs = [t.to(DEV) for t in synthetic_inputs(cfg, batch, seed=11)]
    adj, node, flags, sigmas, sc_adj, sc_node = inputs
    outs = []
    for no_fuse in ("0", "1"):
        os.environ["DSG_NO_FUSED_MLP"] = no_fuse
        try:
            model, _ = build(cfg)
            with torch.no_grad():
                outs.append(model(adj, node, flags, sigmas.log() / 4, sc_adj, sc_node))
        finally:
            os.environ.pop("DSG_NO_FUSED_MLP", None)
    (fa, fn), (ua, un) = outs
    assert torch.isfinite(fa).all() and torch.isfinite(fn).all()
    assert rel(fa, ua) < 3e-3 and rel(fn, un) < 3e-3, (rel(fa, ua), rel(fn, un))


@pytest.mark.parametrize("name,batch", [("vg", 3), ("vg", 40)])
def test_block_tail_matches_unfused_schedule(name, batch):
    """The fused proj + residual + LN2 + MLP kernel (C = 96 blocks) against the proj GEMM, LayerNorm and fused-MLP
    launches it replaces: same bf16 roundings of the LN output and of the hidden activation; the fp32 accumulation
    order and the (pivot-shifted one-pass) LayerNorm statistics differ, which flips bf16 roundings: measured 5.5e-3
    on the stress weights, with both schedules at the same 0.97e-2 from the oracle (test_forward_stage_by_stage)."""
    cfg = CONFIGS[name]
    inputs = [t.to(DEV) for t in synthetic_inputs(cfg, batch, seed=13)]
    adj, node, flags, sigmas, sc_adj, sc_node = inputs
    outs = []
    for no_tail in ("0", "1"):
        os.environ["DSG_NO_TAIL"] = no_tail
        try:
            model, _ = build(cfg)
            with torch.no_grad():
                outs.append(model(adj, node, flags, sigmas.log() / 4, sc_adj, sc_node))
        finally:
            os.environ.pop("DSG_NO_TAIL", None)
    (fa, fn), (ua, un) = outs
    assert torch.isfinite(fa).all() and torch.isfinite(fn).all()
    assert rel(fa, ua) < 8e-3 and rel(fn, un) < 8e-3, (rel(fa, ua), rel(fn, un))


@pytest.mark.parametrize("name,batch", [("vg", 3), ("tiny", 5)])
def test_block_head_matches_unfused_schedule(name, batch):
    """The fused FiLM + LN1 + qkv kernel (C = 96 blocks) against the FiLM/LayerNorm row kernel + qkv GEMM it replaces:
    same fp32 FiLM/SiLU/LayerNorm expressions and the same bf16 rounding of the LayerNorm output; only the fp32
    accumulation order of the statistics and of the GEMM differs.  tiny has 256-token samples and per-sample sigmas,
    i.e. per-row FiLM rows inside one 128-token tile sequence."""
    cfg = CONFIGS[name]
    inputs = [t.to(DEV) for t in synthetic_inputs(cfg, batch, seed=17)]
    adj, node, flags, sigmas, sc_adj, sc_node = inputs
    labels = (sigmas * torch.linspace(0.5, 2.0, batch, device=DEV)).log() / 4
    outs = []
    for no_head in ("0", "1"):
        os.environ["DSG_NO_HEAD"] = no_head
        try:
            model, _ = build(cfg)
            with torch.no_grad():
                outs.append(model(adj, node, flags, labels, sc_adj, sc_node))
        finally:
            os.environ.pop("DSG_NO_HEAD", None)
    (fa, fn), (ua, un) = outs
    assert torch.isfinite(fa).all() and torch.isfinite(fn).all()
    assert rel(fa, ua) < 8e-3 and rel(fn, un) < 8e-3, (rel(fa, ua), rel(fn, un))


def test_forward_uniform_vs_per_sample_sigma():
    cfg = CONFIGS["tiny"]
    model, _ = build(cfg)
    adj, node, flags, sigmas, sc_adj, sc_node = [t.to(DEV) for t in synthetic_inputs(cfg, 4, seed=3)]
    lab = torch.tensor(0.3, device=DEV)
    with torch.no_grad():
        a1, n1 = model(adj, node, flags, lab.view(-1).expand(4), sc_adj, sc_node)
        a2, n2 = model(adj, node, flags, lab.repeat(4), sc_adj, sc_node)
    assert torch.equal(a1, a2) and torch.equal(n1, n2)


def test_precond_matches_golden(golden_dir):
    cfg = CONFIGS["tiny"]
    g = np.load(os.path.join(golden_dir, "precond_tiny.npz"))
    net, _ = build(cfg)
    model = NodeAdjPrecond(precond="edm", model=net, self_condition=True, symmetric_noise=False).eval()
    adj, node, flags, _, _, _ = [t.to(DEV) for t in synthetic_inputs(cfg, 3, seed=7)]
    np.random.seed(5)
    sa = sn = None
    with torch.no_grad():
        for k, s in enumerate((40.0, 3.0, 0.4, 0.01)):
            sa, sn = model(adj * s, node * s, flags, torch.full((3,), s, device=DEV), sa, sn)
            c_out = s * 0.5 / (s * s + 0.25) ** 0.5
            for got, key in ((sa, f"adj_{k}"), (sn, f"node_{k}")):
                want = torch.from_numpy(g[key])
                err = float((got.cpu() - want).abs().max())
                assert err < F_TOL * c_out * 4 + 1e-5, (key, err)
    assert model.raw_passes == 4 + int((g["coins"] < 0.5).sum())


def _replay(noise_log):
    it = iter(noise_log)
    return lambda shape: next(it).reshape(shape)


def test_sampler_matches_oracle_with_replayed_noise():
    """Same init noise, same per-step noise, same coin flips: native sampler vs the oracle loop driving the
    oracle network (fp32 CPU).  8 steps on the tiny geometry."""
    cfg = CONFIGS["tiny"]
    net, sd = build(cfg)
    model = NodeAdjPrecond(precond="edm", model=net, self_condition=True, symmetric_noise=False).eval()
    _, _, flags, _, _, _ = synthetic_inputs(cfg, 4, seed=7)
    steps = 8
    sampler = NodeAdjEDMSampler(num_steps=steps, clip_samples=True, clip_samples_min=-1.0, clip_samples_max=1.0,
                                clip_samples_scope="x_0", dev=DEV, objective="edm", self_condition=True,
                                symmetric_noise=False)
    # record the noise the native run draws (CPU init + device per-step), then replay it through the oracle
    log = []
    real_randn, real_like = torch.randn, torch.randn_like

    def rec_randn(*a, **k):
        t = real_randn(*a, **k)
        log.append(t.detach().cpu().clone())
        return t

    def rec_like(x, **k):
        t = real_like(x, **k)
        log.append(t.detach().cpu().clone())
        return t

    torch.manual_seed(21)
    np.random.seed(21)
    torch.randn, torch.randn_like = rec_randn, rec_like
    try:
        a, n = sampler.sample(model=model, node_flags=flags.to(DEV), num_node_chan=cfg["c_n"], num_edge_chan=cfg["c_e"])
    finally:
        torch.randn, torch.randn_like = real_randn, real_like
    assert len(log) == 2 + 2 * steps
    np.random.seed(21)
    onet = oracle_net(cfg, sd)
    omodel = lambda aa, nn_, f, sig, sa, sn: O.precond_forward(onet, aa, nn_, f, sig, sa, sn, coin=np.random.rand)
    with torch.no_grad():
        oa, on = E.sample(omodel, flags, cfg["c_e"], cfg["c_n"], num_steps=steps, normal=_replay(log))
    # decode rule of the reference (runner/sampler/sampler_node_adj.py:222-285): clamp -> sign -> bits -> int;
    # boxes are the last 4 node channels, mapped to [0, 1] by x * 0.5 + 0.5
    valid_pair = (flags[:, :, None] & flags[:, None, :])
    nb = cfg["c_n"] - 4
    edge_g, edge_o = E.decode_bits(a.permute(0, 2, 3, 1), 2 ** cfg["c_e"]), E.decode_bits(oa.permute(0, 2, 3, 1), 2 ** cfg["c_e"])
    node_g, node_o = E.decode_bits(n[..., :nb], 2 ** nb), E.decode_bits(on[..., :nb], 2 ** nb)
    edge_agree = float((edge_g == edge_o)[valid_pair].float().mean())
    node_agree = float((node_g == node_o)[flags].float().mean())
    box_err = ((n[..., nb:] - on[..., nb:]).abs() * 0.5)[flags]
    stats = dict(rel_adj=rel(a, oa), rel_node=rel(n, on), max_adj=float((a - oa).abs().max()),
                 max_node=float((n - on).abs().max()), edge_agree=edge_agree, node_agree=node_agree,
                 box_within_1e2=float((box_err <= 1e-2).float().mean()), passes=sampler.last_raw_passes)
    print("SAMPLER_PARITY", stats)
    assert stats["rel_adj"] < 3e-2 and stats["rel_node"] < 3e-2, stats
    # with random-init weights many final values sit near the sign threshold (of the 204 valid node pairs, 3 to 8
    # flip depending on which bf16 schedule runs, at an unchanged 2.2e-2 state error), so the stated rates are:
    # >= 95 % overall after 8 steps, and 100 % wherever the reference value is further than the state tolerance
    # (0.06) from the decision threshold; boxes within 1e-2 (on the [0, 1] scale) for >= 99 %
    assert edge_agree >= 0.95 and node_agree >= 0.95, stats
    far_e = (oa.abs() > 0.06).permute(0, 2, 3, 1).all(-1) & valid_pair
    far_n = (on[..., :nb].abs() > 0.06).all(-1) & flags
    assert bool((edge_g == edge_o)[far_e].all()) and bool((node_g == node_o)[far_n].all()), stats
    assert stats["box_within_1e2"] >= 0.99, stats
    assert sampler.last_raw_passes >= 2 * steps - 1


def test_sampler_known_answer(golden_dir):
    """The reference's own KAT (sanity_check_gt_*, edm.py:372-377): bit-identical to the reference run."""
    g = np.load(os.path.join(golden_dir, "sampler_tiny.npz"))
    cfg = CONFIGS["tiny"]
    _, _, flags, _, _, _ = synthetic_inputs(cfg, 3, seed=7)
    flags = flags[:2]
    gt_a, gt_n = torch.from_numpy(g["kat_gt_adjs"]), torch.from_numpy(g["kat_gt_nodes"])
    sampler = NodeAdjEDMSampler(num_steps=8, clip_samples=True, clip_samples_min=-1.0, clip_samples_max=1.0,
                                clip_samples_scope="x_0", dev=DEV, objective="edm", self_condition=True,
                                symmetric_noise=False)

    class _NoModel:
        round_sigma = staticmethod(torch.as_tensor)

    # the reference draws the per-step noise on its device (CPU for the golden run): replay the CPU stream
    torch.manual_seed(12)
    init_a = torch.randn(2, cfg["c_e"], cfg["img"], cfg["img"])
    init_n = torch.randn(2, cfg["img"], cfg["c_n"])
    eps = []
    for _ in range(8):
        eps.append(torch.randn(2, cfg["c_e"], cfg["img"], cfg["img"]))
        eps.append(torch.randn(2, cfg["img"], cfg["c_n"]))
    it = iter(eps)
    real_like = torch.randn_like
    torch.randn_like = lambda x, **k: next(it).to(x.device)
    try:
        a, n = sampler.sample(model=_NoModel(), node_flags=flags.to(DEV), init_adjs=O.mask_pairs(init_a, flags).to(DEV),
                              init_nodes=O.mask_rows(init_n, flags).to(DEV), sanity_check_gt_adjs=gt_a.to(DEV),
                              sanity_check_gt_nodes=gt_n.to(DEV), num_node_chan=cfg["c_n"], num_edge_chan=cfg["c_e"])
    finally:
        torch.randn_like = real_like
    np.testing.assert_array_equal(a.numpy(), g["kat_adjs"])
    np.testing.assert_array_equal(n.numpy(), g["kat_nodes"])
    assert float((a - gt_a).abs().max()) < 1e-6 and float((n - gt_n).abs().max()) < 1e-6


def test_sampler_fused_noise_is_bit_identical():
    """The sampler with the per-step noise drawn inside the pre-step kernel == the sampler with the reference's two
    torch.randn_like launches per step: same seeds, bit-identical samples, same generator state afterwards."""
    cfg = CONFIGS["tiny"]
    net, _ = build(cfg)
    model = NodeAdjPrecond(precond="edm", model=net, self_condition=True, symmetric_noise=False).eval()
    _, _, flags, _, _, _ = synthetic_inputs(cfg, 6, seed=9)
    outs = []
    for fused in (True, False):
        sampler = NodeAdjEDMSampler(num_steps=6, clip_samples=True, clip_samples_min=-1.0, clip_samples_max=1.0,
                                    clip_samples_scope="x_0", dev=DEV, objective="edm", self_condition=True,
                                    symmetric_noise=False)
        sampler.fused_noise = fused
        torch.manual_seed(5)
        torch.cuda.manual_seed(5)
        np.random.seed(5)
        a, n = sampler.sample(model=model, node_flags=flags.to(DEV), num_node_chan=cfg["c_n"], num_edge_chan=cfg["c_e"])
        outs.append((a, n, torch.randn(100, device=DEV).cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][2], outs[1][2])


def test_no_cpu_fallback():
    cfg = CONFIGS["tiny"]
    m = DiffuseSG(img_size=cfg["img"], in_chans=in_chans(cfg), patch_size=1, embed_dim=96, depths=cfg["depths"],
                  num_heads=[3, 6, 12, 24], window_size=cfg["window"], mlp_ratio=4., drop_rate=0., attn_drop_rate=0.,
                  drop_path_rate=0., self_condition=True, symmetric_noise=False, out_chans_adj=cfg["c_e"],
                  out_chans_node=cfg["c_n"]).eval()
    adj, node, flags, sigmas, _, _ = synthetic_inputs(cfg, 2, seed=7)
    with pytest.raises(native.NativeError):
        with torch.no_grad():
            m(adj, node, flags, sigmas.log() / 4)


def test_sampler_graphs_are_bit_identical():
    """One CUDA-graph launch per step == the eager launch sequence: same samples, same interim snapshots, same
    generator state afterwards, same pass count; also with every 3rd step taken eagerly (the bench's profiling mode)
    and when the same sampler is called twice (buffers of the plan are reused)."""
    cfg = CONFIGS["tiny"]
    net, _ = build(cfg)
    model = NodeAdjPrecond(precond="edm", model=net, self_condition=True, symmetric_noise=False).eval()
    _, _, flags, _, _, _ = synthetic_inputs(cfg, 6, seed=9)
    outs = []
    for graphs, eager_every in ((False, 0), (True, 0), (True, 3), (True, 0)):
        sampler = NodeAdjEDMSampler(num_steps=10, clip_samples=True, clip_samples_min=-1.0, clip_samples_max=1.0,
                                    clip_samples_scope="x_0", dev=DEV, objective="edm", self_condition=True,
                                    symmetric_noise=False)
        sampler.use_graphs, sampler.eager_every = graphs, eager_every
        for rep in range(2):
            torch.manual_seed(5)
            torch.cuda.manual_seed(5)
            np.random.seed(5)
            p0 = model.raw_passes
            a, n, a_ls, n_ls = sampler.sample(model=model, node_flags=flags.to(DEV), flag_interim_adjs=True,
                                              max_num_interim_adjs=4, num_node_chan=cfg["c_n"], num_edge_chan=cfg["c_e"])
            outs.append((a, n, a_ls, n_ls, torch.randn(100, device=DEV).cpu(), torch.tensor(model.raw_passes - p0)))
        assert (sampler._plan is not None) == graphs
    for o in outs[1:]:
        for x, y in zip(outs[0], o):
            assert torch.equal(x, y)


def test_weights_written_through_data_are_noticed():
    """ema_pytorch updates the EMA copy with `.data.lerp_()` / `.data.copy_()`, which moves no autograd version
    counter: the native arena must still follow.  Forward after an in-place `.data` write == forward of
    a freshly built model with the same weights; and the sampler (frozen inside its loop) re-probes on every call."""
    cfg = CONFIGS["tiny"]
    net, sd = build(cfg)
    adj, node, flags, sigmas, sc_adj, sc_node = [t.to(DEV) for t in synthetic_inputs(cfg, 3, seed=7)]
    labels = sigmas.log() / 4
    with torch.no_grad():
        a0, n0 = net(adj, node, flags, labels, sc_adj, sc_node)
        versions = [p._version for p in net.parameters()]
        for p in net.parameters():
            p.data.mul_(1.25)
        assert versions == [p._version for p in net.parameters()]     # the counters did not move
        a1, n1 = net(adj, node, flags, labels, sc_adj, sc_node)
    fresh = DiffuseSG(img_size=cfg["img"], in_chans=in_chans(cfg), patch_size=1, embed_dim=cfg["embed"],
                      depths=cfg["depths"], num_heads=[3, 6, 12, 24], window_size=cfg["window"], mlp_ratio=4.,
                      drop_rate=0., attn_drop_rate=0., drop_path_rate=0.0, self_condition=True, symmetric_noise=False,
                      out_chans_adj=cfg["c_e"], out_chans_node=cfg["c_n"])
    fresh.load_state_dict({k: (v * 1.25 if v.dtype == torch.float32 and "attn_mask" not in k else v) for k, v in sd.items()})
    fresh = fresh.to(DEV).eval()
    with torch.no_grad():
        a2, n2 = fresh(adj, node, flags, labels, sc_adj, sc_node)
    assert torch.equal(a1, a2) and torch.equal(n1, n2)
    assert not torch.equal(a0, a1)
    # sampler: weights changed between two sample() calls through .data only
    model = NodeAdjPrecond(precond="edm", model=net, self_condition=True, symmetric_noise=False).eval()
    sampler = NodeAdjEDMSampler(num_steps=3, clip_samples=True, clip_samples_min=-1.0, clip_samples_max=1.0,
                                clip_samples_scope="x_0", dev=DEV, objective="edm", self_condition=True,
                                symmetric_noise=False)
    res = []
    for scale in (1.0, 0.5, 2.0):
        with torch.no_grad():
            for p in net.parameters():
                p.data.mul_(scale)
        torch.manual_seed(1); torch.cuda.manual_seed(1); np.random.seed(1)
        res.append(sampler.sample(model=model, node_flags=flags, num_node_chan=cfg["c_n"], num_edge_chan=cfg["c_e"]))
    assert not torch.equal(res[0][0], res[1][0])
    assert torch.equal(res[0][0], res[2][0]) and torch.equal(res[0][1], res[2][1])   # x0.5 then x2: back to the start


def test_sampler_accepts_any_flag_dtype():
    """mask_adjs / mask_nodes of the reference take any flag dtype (logical_not); float and int64 flags must mask the
    same rows as bool flags."""
    cfg = CONFIGS["tiny"]
    net, _ = build(cfg)
    model = NodeAdjPrecond(precond="edm", model=net, self_condition=True, symmetric_noise=False).eval()
    _, _, flags, _, _, _ = synthetic_inputs(cfg, 4, seed=9)
    sampler = NodeAdjEDMSampler(num_steps=2, clip_samples=True, clip_samples_min=-1.0, clip_samples_max=1.0,
                                clip_samples_scope="x_0", dev=DEV, objective="edm", self_condition=True,
                                symmetric_noise=False)
    res = []
    for f in (flags, flags.float(), flags.long()):
        torch.manual_seed(2); torch.cuda.manual_seed(2); np.random.seed(2)
        res.append(sampler.sample(model=model, node_flags=f.to(DEV), num_node_chan=cfg["c_n"], num_edge_chan=cfg["c_e"]))
    for r in res[1:]:
        assert torch.equal(r[0], res[0][0]) and torch.equal(r[1], res[0][1])


def _flags_with_counts(cfg, counts):
    return torch.arange(cfg["img"])[None, :] < torch.tensor(counts)[:, None]


@pytest.mark.parametrize("name,counts", [("vg", [2, 9, 16, 17, 33, 48, 62, 5]), ("tiny", [1, 4, 5, 8, 14, 3]),
                                         ("coco", [2, 10, 11, 33, 21, 30]), ("vg", [20]), ("vg", [40, 0]),
                                         ("vg", [30] * 5), ("coco", [7])])
def test_padding_skipping_matches_dense(name, counts):
    """the compact schedule (leading un-shifted stages computed only on each sample's corner that can hold
    valid nodes, the all-padding region represented by one phantom token) against the dense schedule on the same
    inputs.  Equal in exact arithmetic; here both sides are bf16 tensor-core runs whose only difference is WHERE the
    padding tokens were computed, so the gap must stay far below the parity tolerance (2e-2) - and masked outputs
    must be exactly zero in both."""
    cfg = CONFIGS[name]
    net, sd = build(cfg)
    b = len(counts)
    adj, node, _, _, sc_adj, sc_node = synthetic_inputs(cfg, b, seed=11)
    flags = _flags_with_counts(cfg, counts)
    if b > 2:
        flags[0, 0] = False                      # node flags need not be a prefix: holes, and ...
        if counts[1] + 2 < cfg["img"]:
            flags[1, counts[1] + 1] = True       # ... a valid node beyond a gap (the kept corner follows the LAST valid node)
    pair = (flags[:, None, :, None] & flags[:, None, None, :]).float()
    adj, sc_adj = adj.abs().clamp_min(0.1) * adj.sign() * pair, sc_adj * pair   # re-mask for these flags
    node, sc_node = node * flags[:, :, None], sc_node * flags[:, :, None]
    stages, granule = net._native(DEV).skip_info()
    assert stages >= 1, (name, stages, granule)
    plan = net.make_skip_plan(flags.to(DEV), force=b <= 2)   # tiny batches: the phantom may outweigh what is skipped
    assert plan is not None and (plan.kept_fraction < 1.0 or b <= 2)
    sig = torch.full((1,), 1.5, device=DEV).expand(b)       # one shared noise level, as in sampling
    args = [t.to(DEV) for t in (adj, node, flags)]
    with torch.no_grad():
        n0 = native.launch_count()
        da, dn = net.denoise(args[0], args[1], args[2], sig, sc_adj.to(DEV), sc_node.to(DEV))
        n1 = native.launch_count()
        with net.skipping(plan):
            sa, sn = net.denoise(args[0], args[1], args[2], sig, sc_adj.to(DEV), sc_node.to(DEV))
            # the compact schedule ran: its geometry-aware kernels launch once per bucket
            assert native.launch_count() - n1 > n1 - n0 or len(plan.counts) == 1
            sa2, sn2 = net.denoise(args[0], args[1], args[2], sig, None, None)
        da2, dn2 = net.denoise(args[0], args[1], args[2], sig, None, None)
        # reference: oracle D on the CPU
        oa, on = O.precond_forward(oracle_net(cfg, sd), adj, node, flags, torch.full((b,), 1.5), sc_adj, sc_node, coin=lambda: 1.0)
    stats = dict(name=name, kept=plan.kept_fraction, skip_vs_dense=(rel(sa, da), rel(sn, dn)), nosc=(rel(sa2, da2), rel(sn2, dn2)),
                 skip_vs_oracle=(rel(sa, oa), rel(sn, on)), dense_vs_oracle=(rel(da, oa), rel(dn, on)))
    print("SKIP_PARITY", stats)
    assert max(stats["skip_vs_dense"] + stats["nosc"]) < 3e-3, stats
    assert max(stats["skip_vs_oracle"]) < max(1.5 * max(stats["dense_vs_oracle"]), 1e-3), stats
    invalid = ~(flags[:, None, :, None] & flags[:, None, None, :]).expand_as(adj)
    assert float(sa.cpu()[invalid].abs().sum()) == 0.0 and float(sn.cpu()[~flags].abs().sum()) == 0.0
    # a batch without padding keeps the dense schedule
    full = _flags_with_counts(cfg, [cfg["img"]] * 3)
    assert net.make_skip_plan(full.to(DEV)) is None


@pytest.mark.parametrize("name", ["vg", "coco"])
def test_padding_skipping_precond_matches_golden(name, golden_dir):
    """NodeAdjPrecond.forward of the unmodified reference (golden) against the native call with padding skipping
    active: same tolerance as the dense path (test_precond_matches_golden)."""
    cfg = CONFIGS[name]
    g = np.load(os.path.join(golden_dir, f"precond_{name}.npz"))
    net, _ = build(cfg)
    model = NodeAdjPrecond(precond="edm", model=net, self_condition=True, symmetric_noise=False).eval()
    adj, node, flags, _, _, _ = [t.to(DEV) for t in synthetic_inputs(cfg, 2, seed=7)]
    plan = net.make_skip_plan(flags, force=True)    # batch 2: the phantom may cost more rows than are skipped
    assert plan is not None
    np.random.seed(5)
    sa = sn = None
    with torch.no_grad(), net.skipping(plan):
        for k, s in enumerate((40.0, 3.0, 0.4, 0.01)):
            sa, sn = model(adj * s, node * s, flags, torch.full((1,), s, device=DEV).expand(2), sa, sn)
            c_out = s * 0.5 / (s * s + 0.25) ** 0.5
            for got, key in ((sa, f"adj_{k}"), (sn, f"node_{k}")):
                want = torch.from_numpy(g[key])
                err = float((got.cpu() - want).abs().max())
                assert err < F_TOL * c_out * 4 + 1e-5, (key, err)


def test_sampler_graphs_follow_changing_flags():
    """One sampler, alternating batches of node flags (different padding plans, same batch size): the captured graphs are
    keyed by the plan geometry and read the refreshed tables; results equal the eager dense loop bit for bit."""
    cfg = CONFIGS["vg"]
    net, _ = build(cfg, stress=False)
    model = NodeAdjPrecond(precond="edm", model=net, self_condition=True, symmetric_noise=False).eval()
    flag_sets = [_flags_with_counts(cfg, [5, 17, 33, 62, 20, 48]), _flags_with_counts(cfg, [30, 31, 2, 9, 64, 16]),
                 _flags_with_counts(cfg, [5, 17, 33, 62, 20, 48]), _flags_with_counts(cfg, [64] * 6)]
    outs = {}
    for mode in ("graphs+skip", "eager-dense"):
        sampler = NodeAdjEDMSampler(num_steps=4, clip_samples=True, clip_samples_min=-1.0, clip_samples_max=1.0,
                                    clip_samples_scope="x_0", dev=DEV, objective="edm", self_condition=True,
                                    symmetric_noise=False)
        sampler.use_graphs = mode == "graphs+skip"
        sampler.skip_padding = mode == "graphs+skip"
        res = []
        for k, flags in enumerate(flag_sets):
            torch.manual_seed(40 + k); torch.cuda.manual_seed(40 + k); np.random.seed(40 + k)
            res.append(sampler.sample(model=model, node_flags=flags.to(DEV), num_node_chan=cfg["c_n"], num_edge_chan=cfg["c_e"]))
        outs[mode] = res
    for (a0, n0), (a1, n1) in zip(outs["graphs+skip"], outs["eager-dense"]):
        assert torch.equal(a0, a1) and torch.equal(n0, n1)


@pytest.mark.parametrize("name,batch", [("tiny", 3), ("vg", 2), ("coco", 2)])
def test_kernels_ignore_shared_memory_leftovers(name, batch):
    """Every launch followed by a kernel that fills all shared memory with NaN / Inf patterns (what NCCL or any foreign
    kernel may leave behind on an SM): the forward (shared and per-sample noise levels, with and without
    self-conditioning) must not change by a bit.  Regression test for node_proj_kernel, which multiplied its zero padding
    weights with shared-memory slots it never wrote (0 x NaN): NaN samples on the first pass of a 2-process run."""
    cfg = CONFIGS[name]
    model, _ = build(cfg)
    adj, node, flags, sigmas, sc_adj, sc_node = [t.to(DEV) for t in synthetic_inputs(cfg, batch, seed=7)]
    per_sample = (sigmas * torch.linspace(0.1, 3.0, batch, device=DEV)).log() / 4
    cases = [(sigmas.log() / 4, sc_adj, sc_node), (per_sample, None, None), ((sigmas[:1].log() / 4).expand(batch), None, None)]
    lib = native.lib()
    with torch.no_grad():
        want = [model(adj, node, flags, lab, sa, sn) for lab, sa, sn in cases]
        for pattern in (0x7fc00000, 0xff800000, 0x7f7fffff):
            lib.dsg_debug_set_smem_poison(pattern, 1)
            try:
                got = [model(adj, node, flags, lab, sa, sn) for lab, sa, sn in cases]
                torch.cuda.synchronize()
            finally:
                lib.dsg_debug_set_smem_poison(0, 0)
            for (wa, wn), (ga, gn) in zip(want, got):
                assert torch.isfinite(ga).all() and torch.isfinite(gn).all(), hex(pattern)
                assert torch.equal(wa, ga) and torch.equal(wn, gn), hex(pattern)


@pytest.mark.parametrize("name,batch", [("tiny", 6), ("vg", 4), ("coco", 4)])
def test_sampler_ignores_shared_memory_leftovers(name, batch):
    """The whole eager sampler path (EDM step kernels with in-kernel noise, padding-skipping schedule, fused decode of the
    last step) with all shared memory overwritten by a NaN pattern after every launch: bit-identical samples and classes."""
    cfg = CONFIGS[name]
    net, _ = build(cfg)
    model = NodeAdjPrecond(precond="edm", model=net, self_condition=True, symmetric_noise=False).eval()
    _, _, flags, _, _, _ = synthetic_inputs(cfg, batch, seed=9)
    outs = []
    for poison in (0, 1):
        sampler = NodeAdjEDMSampler(num_steps=4, clip_samples=True, clip_samples_min=-1.0, clip_samples_max=1.0,
                                    clip_samples_scope="x_0", dev=DEV, objective="edm", self_condition=True,
                                    symmetric_noise=False)
        sampler.use_graphs = False            # the poison kernel runs on the legacy default stream between eager launches
        torch.manual_seed(5)
        torch.cuda.manual_seed(5)
        np.random.seed(5)
        native.lib().dsg_debug_set_smem_poison(0x7fc00000, poison)
        try:
            out = sampler.sample_decoded(model, flags.to(DEV), 7, 150, num_node_chan=cfg["c_n"], num_edge_chan=cfg["c_e"],
                                         return_state=True)
            torch.cuda.synchronize()
        finally:
            native.lib().dsg_debug_set_smem_poison(0, 0)
        outs.append(out)
    for x, y in zip(*outs):
        assert torch.isfinite(y.float()).all() and torch.equal(x, y)
