"""Kernel-level parity on the B200 (run with -m gpu): every check goes through the C ABI of libdsg_b200.so."""
import numpy as np
import pytest
import torch

from diffusesg_b200 import native
from diffusesg_b200.model.diffusesg.geometry import relative_position_index, shifted_window_mask
from oracle import edm_oracle as E
from oracle.denoiser_oracle import mask_pairs, mask_rows

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


# ---------------------------------------------------------------------------------------------------------
# tcgen05 GEMM
# ---------------------------------------------------------------------------------------------------------
GEMM_SHAPES = [
    # (M, N, K): tails in M (TMA zero fill + predicated stores), K = 96 (half-empty second k-block), both tile widths
    (128, 96, 64), (128, 192, 64), (256, 96, 96), (300, 288, 96), (1000, 384, 192), (4096, 1152, 384),
    (777, 768, 3072), (512, 1536, 1536), (48, 96, 384), (20000, 96, 96), (33000, 576, 192),
    # CTA-pair path (cta_group::2, K >= 768, M >= 256): row tails inside the second CTA's half, both tile widths
    (256, 96, 768), (300, 384, 768), (3000, 2304, 768), (1111, 96, 1536),
    # 256-wide tiles (N % 256 == 0, K >= 384, enough tiles to fill the waves): single CTA and CTA pairs, row tails
    (20000, 1536, 384), (19999, 768, 384), (40000, 768, 768), (33333, 1536, 768),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi", [native.EPI_BF16, native.EPI_GELU_BF16, native.EPI_RES_F32, native.EPI_F32])
def test_gemm_matches_torch(M, N, K, epi):
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N * 3 + K + epi)
    a = (torch.randn(M, K, device=DEV, generator=g)).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV, generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV, generator=g)
    res = torch.randn(M, N, device=DEV, generator=g) if epi == native.EPI_RES_F32 else None
    want = a.float() @ w.float().t() + bias
    if epi == native.EPI_GELU_BF16:
        want = torch.nn.functional.gelu(want)
    if res is not None:
        want = want + res
    got = native.gemm_bf16(a, w, bias, res, epi).float()
    torch.cuda.synchronize()
    err = (got - want).abs()
    tol = 2e-2 if epi in (native.EPI_BF16, native.EPI_GELU_BF16) else 2e-3   # bf16 output rounding vs fp32 output
    bad = (err > tol * (1 + want.abs())).nonzero()
    assert bad.numel() == 0, (f"{bad.shape[0]} bad of {M * N}; first {bad[:5].tolist()}; max err {float(err.max())}; "
                              f"bad rows {sorted(set(bad[:, 0].tolist()))[:10]} cols {sorted(set(bad[:, 1].tolist()))[:10]}")
    assert _rel(got, want) < (5e-3 if tol > 1e-2 else 1e-5 + 2e-6 * K ** 0.5)


def test_gemm_residual_in_place():
    M, N, K = 1000, 192, 384
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).to(torch.bfloat16)
    x = torch.randn(M, N, device=DEV)
    want = x + a.float() @ w.float().t()
    native.check(native.lib().dsg_gemm_bf16(a.data_ptr(), w.data_ptr(), None, x.data_ptr(), x.data_ptr(), M, N, K,
                                            native.EPI_RES_F32, native.stream_ptr()), "gemm")
    assert _rel(x, want) < 1e-4


@pytest.mark.parametrize("M,C", [(128, 192), (1000, 192), (4096, 384), (777, 384), (40000, 384), (33333, 192)])
def test_proj_ln_matches_torch(M, C):
    """Fused attention projection + residual + LayerNorm2 (C = 192 / 384) against fp32 torch on the same bf16 operands."""
    g = torch.Generator(device=DEV).manual_seed(M + C)
    att = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    w = (torch.randn(C, C, device=DEV, generator=g) / C ** 0.5).to(torch.bfloat16)
    bias = torch.randn(C, device=DEV, generator=g) * 0.1
    gamma = 1 + 0.2 * torch.randn(C, device=DEV, generator=g)
    beta = 0.1 * torch.randn(C, device=DEV, generator=g)
    x = torch.randn(M, C, device=DEV, generator=g) * 2 + 0.7   # a non-zero row mean exercises the pivot
    want_x = x + (att.float() @ w.float().t() + bias)
    want_y = torch.nn.functional.layer_norm(want_x, (C,), gamma, beta, 1e-5)
    y = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    native.check(native.lib().dsg_proj_ln(att.data_ptr(), w.data_ptr(), bias.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                          x.data_ptr(), y.data_ptr(), M, C, native.stream_ptr()), "proj_ln")
    torch.cuda.synchronize()
    assert _rel(x, want_x) < 1e-5, _rel(x, want_x)
    assert float((y.float() - want_y).abs().max()) < 3e-2 and _rel(y.float(), want_y) < 4e-3   # bf16 output rounding


def test_gemm_rejects_bad_shapes():
    a = torch.zeros(128, 64, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(100, 64, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(native.NativeError):
        native.gemm_bf16(a, w)


# ---------------------------------------------------------------------------------------------------------
# window attention
# ---------------------------------------------------------------------------------------------------------
def _attention_reference(qkv, bias, mask, batch, res, w, shift, heads):
    c = heads * 32
    x = qkv.float().view(batch, res, res, 3 * c)
    if shift:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    nw = res // w
    t = w * w
    xw = x.view(batch, nw, w, nw, w, 3 * c).permute(0, 1, 3, 2, 4, 5).reshape(batch * nw * nw, t, 3, heads, 32)
    q, k, v = xw.permute(2, 0, 3, 1, 4)
    att = q @ k.transpose(-1, -2) + bias[None]
    if mask is not None:
        att = (att.view(batch, nw * nw, heads, t, t) + mask[None, :, None]).view(-1, heads, t, t)
    out = (att.softmax(-1) @ v).transpose(1, 2).reshape(batch, nw, nw, w, w, c).permute(0, 1, 3, 2, 4, 5)
    out = out.reshape(batch, res, res, c)
    if shift:
        out = torch.roll(out, shifts=(shift, shift), dims=(1, 2))
    return out.reshape(batch * res * res, c)


@pytest.mark.parametrize("batch,res,w,shift,heads", [
    (3, 16, 4, 0, 3), (3, 8, 4, 2, 6), (2, 64, 8, 0, 3), (5, 16, 8, 4, 12), (2, 8, 8, 0, 24),
    (2, 20, 10, 5, 6), (3, 10, 10, 0, 12), (1, 32, 16, 8, 6), (2, 16, 16, 0, 12),
    # un-shifted 8 x 8 windows with an even window count take the tcgen05 kernel (two windows per MMA tile)
    (3, 16, 8, 0, 12), (4, 32, 8, 0, 6), (37, 16, 8, 0, 3), (1, 8, 8, 0, 3),
    # even windows up to 10 x 10, shifted or not, take the quad-box tcgen05 kernel (one window per tile; the COCO-Stuff
    # geometry: 10 x 10 windows at res 40 / 20 / 10, shift 5 at res 20; the VG shifted blocks: 8 x 8, shift 4 at res 16)
    (2, 40, 10, 0, 3), (3, 20, 10, 0, 6), (5, 20, 10, 5, 6), (1, 30, 10, 5, 3), (7, 16, 8, 4, 12), (3, 24, 8, 4, 3),
    (2, 12, 6, 3, 3), (41, 10, 10, 0, 24), (2, 32, 8, 4, 6), (512, 16, 8, 4, 12),
    # 16 x 16 windows (BASELINE config 5): two 128-row query halves per window-head against 256 keys
    (3, 64, 16, 0, 3), (2, 64, 16, 8, 3), (3, 32, 16, 8, 6), (5, 16, 16, 0, 12), (2, 48, 16, 8, 3)])
def test_window_attention_matches_torch(batch, res, w, shift, heads):
    g = torch.Generator(device=DEV).manual_seed(res * 100 + w + shift)
    c = heads * 32
    t = w * w
    qkv = torch.randn(batch * res * res, 3 * c, device=DEV, generator=g)
    qkv[:, :c] *= 32 ** -0.5
    qkv = qkv.to(torch.bfloat16)
    table = torch.randn((2 * w - 1) ** 2, heads, device=DEV, generator=g) * 0.5
    bias = table[relative_position_index(w).to(DEV).reshape(-1)].view(t, t, heads).permute(2, 0, 1).contiguous()
    mask = shifted_window_mask(res, w, shift).to(DEV) if shift else None
    want = _attention_reference(qkv, bias, mask, batch, res, w, shift, heads)
    got = native.window_attention(qkv, bias, mask, batch, res, w, shift, heads).float()
    torch.cuda.synchronize()
    assert torch.isfinite(got).all()
    assert _rel(got, want) < 1e-2, _rel(got, want)   # p is rounded to bf16 before p.v, output stored as bf16
    assert float((got - want).abs().max()) < 5e-2


@pytest.mark.parametrize("batch,res,w,shift,heads", [(3, 20, 10, 5, 6), (2, 32, 16, 8, 3)])
def test_window_attention_arbitrary_mask_takes_the_generic_kernel(batch, res, w, shift, heads):
    """The tcgen05 kernels generate the SW-MSA mask themselves (and the 16 x 16 one looks the bias up by token
    offset); a mask buffer with other values / a bias that is not a function of the offset must still be honoured
    (the launcher checks the buffers and falls back to the kernel that reads them)."""
    g = torch.Generator(device=DEV).manual_seed(5)
    c, t = heads * 32, w * w
    qkv = torch.randn(batch * res * res, 3 * c, device=DEV, generator=g)
    qkv[:, :c] *= 32 ** -0.5
    qkv = qkv.to(torch.bfloat16)
    bias = torch.randn(heads, t, t, device=DEV, generator=g) * 0.3
    mask = shifted_window_mask(res, w, shift).to(DEV) * 0.05 + torch.randn(4, t, t, device=DEV, generator=g) * 0.2
    want = _attention_reference(qkv, bias, mask, batch, res, w, shift, heads)
    got = native.window_attention(qkv, bias, mask.contiguous(), batch, res, w, shift, heads).float()
    assert _rel(got, want) < 1e-2, _rel(got, want)


@pytest.mark.parametrize("b,ce,n,cn", [(3, 3, 16, 5), (8, 6, 64, 12), (512, 6, 64, 12), (5, 3, 40, 12), (1, 1, 4, 1),
                                       (700, 3, 40, 12)])
def test_edm_pre_step_fused_noise_matches_torch(b, ce, n, cn):
    """Noise drawn inside the pre-step kernel (Philox4x32-10 with ATen's launch geometry) == the reference's
    `randn_like(adjs)` then `randn_like(nodes)` on this device, bit for bit, and the torch generator ends in the same
    state (the next torch draw is identical too)."""
    g = torch.Generator().manual_seed(b * 7 + n)
    flags = (torch.arange(n)[None, :] < torch.randint(1, n + 1, (b, 1), generator=g)).to(DEV)
    adj = (torch.randn(b, ce, n, n, generator=g) * 3).to(DEV)
    node = (torch.randn(b, n, cn, generator=g) * 3).to(DEV)
    for seed, burn in ((1234, 0), (99, 3)):
        torch.cuda.manual_seed(seed)
        for _ in range(burn):
            torch.randn(1000, device=DEV)   # start from a non-zero generator offset
        eps_a, eps_n = torch.randn_like(adj), torch.randn_like(node)
        want_a, want_n = native.edm_pre_step(adj, node, eps_a, eps_n, flags, 0.37)
        next_want = torch.randn(4097, device=DEV)
        torch.cuda.manual_seed(seed)
        for _ in range(burn):
            torch.randn(1000, device=DEV)
        got_a, got_n = native.edm_pre_step_fused_noise(adj, node, flags, 0.37)
        next_got = torch.randn(4097, device=DEV)
        torch.cuda.synchronize()
        assert torch.equal(got_a, want_a) and torch.equal(got_n, want_n)
        assert torch.equal(next_got, next_want)


# ---------------------------------------------------------------------------------------------------------
# fused EDM step kernels: bit-exact against the fp32 expressions of the reference sampler
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("b,ce,n,cn", [(3, 3, 16, 5), (8, 6, 64, 12), (5, 3, 40, 12), (1, 1, 4, 1)])
def test_edm_steps_bit_exact(b, ce, n, cn):
    g = torch.Generator().manual_seed(b + n)
    flags = torch.arange(n)[None, :] < torch.randint(1, n + 1, (b, 1), generator=g)
    adj = mask_pairs(torch.randn(b, ce, n, n, generator=g) * 3, flags)
    node = mask_rows(torch.randn(b, n, cn, generator=g) * 3, flags)
    eps_a, eps_n = torch.randn(b, ce, n, n, generator=g), torch.randn(b, n, cn, generator=g)
    d1 = (mask_pairs(torch.randn(b, ce, n, n, generator=g), flags), mask_rows(torch.randn(b, n, cn, generator=g), flags))
    d2 = (mask_pairs(torch.randn(b, ce, n, n, generator=g), flags), mask_rows(torch.randn(b, n, cn, generator=g), flags))
    ts = E.t_steps_fp32(256)
    cu = lambda t: t.to(DEV)
    for i in (0, 40, 130, 254, 255):
        s = E.step_scalars(ts[i], ts[i + 1], 256)
        t_hat, h, t_prime = s["t_hat"], s["h"], s["t_prime"]
        # oracle (CPU fp32 torch)
        a_hat = mask_pairs(adj + s["noise_coef"] * eps_a, flags)
        n_hat = mask_rows(node + s["noise_coef"] * eps_n, flags)
        inv = 1.0 / t_hat
        ka, kn = inv * a_hat - inv * d1[0], inv * n_hat - inv * d1[1]
        pa, pn = a_hat + h * ka, n_hat + h * kn
        if i == 255:
            na, nn_ = mask_pairs(pa, flags), mask_rows(pn, flags)
        else:
            ip = 1.0 / t_prime
            na = mask_pairs(a_hat + h * (0.5 * ka + 0.5 * (ip * pa - ip * d2[0])), flags)
            nn_ = mask_rows(n_hat + h * (0.5 * kn + 0.5 * (ip * pn - ip * d2[1])), flags)
        # device
        ga_hat, gn_hat = native.edm_pre_step(cu(adj), cu(node), cu(eps_a), cu(eps_n), cu(flags), float(s["noise_coef"]))
        np.testing.assert_array_equal(ga_hat.cpu().numpy(), a_hat.numpy())
        np.testing.assert_array_equal(gn_hat.cpu().numpy(), n_hat.numpy())
        ga, gn = native.edm_post_step(ga_hat, gn_hat, (cu(d1[0]), cu(d1[1])), None if i == 255 else (cu(d2[0]), cu(d2[1])),
                                      cu(flags), float(inv), float(h), 0.0 if i == 255 else float(1.0 / t_prime))
        np.testing.assert_array_equal(ga.cpu().numpy(), na.numpy())
        np.testing.assert_array_equal(gn.cpu().numpy(), nn_.numpy())
    sa, sn = native.edm_mask_scale(cu(torch.randn(b, ce, n, n, generator=g)), cu(node), cu(flags), 80.0)
    assert float(sa[~(flags[:, None, :, None] & flags[:, None, None, :]).expand_as(sa).to(DEV)].abs().sum()) == 0.0
    np.testing.assert_array_equal(sn.cpu().numpy(), (node * 80.0).numpy())


# ---------------------------------------------------------------------------------------------------------
# on-device decode of the final sample: integer outputs, bit-exact against the reference rule
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("b,ce,n,cn,n_adj,n_node", [(7, 6, 64, 12, 51, 150), (5, 3, 40, 12, 7, 171), (3, 3, 16, 5, 5, 2)])
def test_decode_samples_bit_exact(b, ce, n, cn, n_adj, n_node):
    g = torch.Generator().manual_seed(n + ce)
    flags = torch.arange(n)[None, :] < torch.randint(1, n + 1, (b, 1), generator=g)
    adj = torch.randn(b, ce, n, n, generator=g) * 1.5          # deliberately NOT masked: the decode masks itself
    node = torch.randn(b, n, cn, generator=g) * 1.5
    adj[0, :, 1, 2] = 0.0                                       # exactly on the threshold -> bit 0
    qa, qn, box = E.decode_samples(adj, node, flags, n_adj, n_node)
    ga, gn, gb = native.decode_samples(adj.to(DEV), node.to(DEV), flags.to(DEV), n_adj, n_node)
    np.testing.assert_array_equal(ga.cpu().numpy(), qa.numpy().astype(np.int32))
    np.testing.assert_array_equal(gn.cpu().numpy(), qn.numpy().astype(np.int32))
    np.testing.assert_array_equal(gb.cpu().numpy(), box.numpy())
    assert int(ga.max()) <= n_adj - 1 and int(gn.max()) <= n_node - 1
    assert int(ga.diagonal(dim1=1, dim2=2).abs().sum()) == 0


@pytest.mark.parametrize("b,ce,n,cn,n_adj,n_node", [(7, 6, 64, 12, 51, 150), (5, 3, 40, 12, 7, 171), (3, 3, 16, 5, 5, 2)])
def test_final_step_decode_fused_bit_exact(b, ce, n, cn, n_adj, n_node):
    """Last Euler step + decode in one kernel == dsg_edm_post_step followed by dsg_decode_samples (and the oracle)."""
    g = torch.Generator().manual_seed(7 * n + ce)
    flags = torch.arange(n)[None, :] < torch.randint(1, n + 1, (b, 1), generator=g)
    xh_a, xh_n = torch.randn(b, ce, n, n, generator=g), torch.randn(b, n, cn, generator=g)
    d_a, d_n = torch.randn(b, ce, n, n, generator=g), torch.randn(b, n, cn, generator=g)
    cu = lambda t: t.to(DEV)
    inv_t, h = 1.0 / 0.0021, -0.0021
    sa, sn = native.edm_post_step(cu(xh_a), cu(xh_n), (cu(d_a), cu(d_n)), None, cu(flags), inv_t, h, 0.0)
    ca, cn_, cb = native.decode_samples(sa, sn, cu(flags), n_adj, n_node)
    fa, fn, qa, qn, qb = native.edm_final_step_decode(cu(xh_a), cu(xh_n), (cu(d_a), cu(d_n)), cu(flags), inv_t, h, n_adj, n_node)
    assert torch.equal(fa, sa) and torch.equal(fn, sn)
    assert torch.equal(qa, ca) and torch.equal(qn, cn_) and torch.equal(qb, cb)
    oa, on, ob = E.decode_samples(sa.cpu(), sn.cpu(), flags, n_adj, n_node)
    np.testing.assert_array_equal(qa.cpu().numpy(), oa.numpy().astype(np.int32))
    np.testing.assert_array_equal(qn.cpu().numpy(), on.numpy().astype(np.int32))
    np.testing.assert_array_equal(qb.cpu().numpy(), ob.numpy())
    na, nn_, qa2, qn2, qb2 = native.edm_final_step_decode(cu(xh_a), cu(xh_n), (cu(d_a), cu(d_n)), cu(flags), inv_t, h, n_adj,
                                                          n_node, want_state=False)
    assert na is None and nn_ is None and torch.equal(qa2, qa) and torch.equal(qn2, qn) and torch.equal(qb2, qb)


def test_training_objective_bit_exact_and_loss(golden_dir):
    """Native noising (dsg_train_noise) and loss (dsg_edm_loss_sums) through the drop-in host classes against the
    golden outputs of the unmodified reference: noising / coefficients bit-exact, loss within 1e-5 relative (fp64
    accumulation here, fp32 pairwise sums there)."""
    import os
    from diffusesg_b200.loss.rainbow_loss import NodeAdjRainbowLoss
    from diffusesg_b200.runner.objectives.edm import NodeAdjEDMObjectiveGenerator
    from diffusesg_b200.utils.synthetic import CONFIGS, synthetic_inputs
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, "train_objective.npz")).items()}
    _, _, flags, _, _, _ = synthetic_inputs(CONFIGS["tiny"], 4, seed=7)
    dev = torch.device("cuda:0")
    # the fused launch on the reference's own draws
    out = native.train_noise(g["clean_a"].to(dev), g["clean_x"].to(dev), g["eps_a"].to(dev), g["eps_x"].to(dev),
                             g["sigmas"].to(dev), flags.to(dev))
    assert torch.equal(out[0].cpu(), g["in_a"]) and torch.equal(out[2].cpu(), g["in_x"])
    pair = flags[:, None, :, None] & flags[:, None, None, :]
    assert torch.equal(out[1].cpu(), torch.where(pair, g["eps_a"] * g["sigmas"].view(-1, 1, 1, 1), torch.zeros(())))
    assert float(out[3].cpu()[~flags].abs().sum()) == 0.0
    # the generator end to end: same RNG consumption order as the reference (device generator here, so only the
    # structure and the masking are checked, not the values)
    gen = NodeAdjEDMObjectiveGenerator("edm", "edm", other_params=None, dev=dev, symmetric_noise=False)
    torch.manual_seed(31)
    in_a, in_x, cond, tgt_a, tgt_x, (c_skip, c_out, c_in, c_noise, sigmas, weights) = gen.get_input_output(
        g["clean_a"].to(dev), g["clean_x"].to(dev), flags.to(dev))
    assert in_a.shape == g["in_a"].shape and in_x.shape == g["in_x"].shape and cond is sigmas
    assert float(in_a.cpu()[~pair.expand_as(in_a)].abs().sum()) == 0.0
    torch.testing.assert_close(weights, (sigmas ** 2 + 0.25) / (sigmas * 0.5) ** 2)
    # loss
    loss = NodeAdjRainbowLoss(edge_loss_weight=1.0, node_loss_weight=0.5, objective="edm")
    for red in ("none", "mean"):
        la, ln = loss(g["pred_a"].to(dev), g["pred_x"].to(dev), g["clean_a"].to(dev), g["clean_x"].to(dev),
                      g["sigmas"].to(dev), None, None, None, None, flags.to(dev), loss_weight=g["weights"].to(dev),
                      reduction=red)
        torch.testing.assert_close(la.cpu(), g[f"loss_adj_{red}"], rtol=1e-5, atol=0)
        torch.testing.assert_close(ln.cpu(), g[f"loss_node_{red}"], rtol=1e-5, atol=0)


@pytest.mark.parametrize("b,ce,n,cn", [(5, 6, 64, 12), (3, 3, 40, 12)])
def test_training_objective_matches_oracle_full_size(b, ce, n, cn):
    from oracle import train_oracle as T
    g = torch.Generator().manual_seed(5)
    flags = torch.zeros(b, n, dtype=torch.bool)
    for i in range(b):
        flags[i, : int(torch.randint(2, n - 1, (1,), generator=g))] = True
    ca, cx = torch.randn(b, ce, n, n, generator=g).sign(), torch.rand(b, n, cn, generator=g) * 2 - 1
    ea, ex = torch.randn(b, ce, n, n, generator=g), torch.randn(b, n, cn, generator=g)
    sig, w = T.training_sigmas_weights(torch.randn(b, generator=g))
    want = T.network_input(ca, cx, flags, sig, ea, ex)
    dev = torch.device("cuda:0")
    got = native.train_noise(ca.to(dev), cx.to(dev), ea.to(dev), ex.to(dev), sig.to(dev), flags.to(dev))
    for a, bb in zip(got, want):
        assert torch.equal(a.cpu(), bb)
    sa, sn = native.edm_loss_sums(got[0], ca.to(dev), got[2], cx.to(dev), w.to(dev), flags.to(dev))
    la, ln = T.regression_loss(want[0], want[2], ca, cx, flags, w, 1.0, 1.0, "none")
    nn_ = flags.sum(-1)
    torch.testing.assert_close(sa.cpu() / nn_ ** 2 / ce, la, rtol=1e-5, atol=0)
    torch.testing.assert_close(sn.cpu() / nn_ / cn, ln, rtol=1e-5, atol=0)


@pytest.mark.parametrize("b,ce,n,cn,red", [(4, 6, 64, 12, "mean"), (3, 3, 40, 12, "none"), (5, 3, 16, 5, "mean")])
def test_loss_backward_matches_torch_autograd(b, ce, n, cn, red):
    """NodeAdjRainbowLoss with predictions that require grad: loss.backward() through the fused backward kernel
    (dsg_edm_loss_sums_backward) against torch autograd of the oracle's restatement of loss/rainbow_loss.py:60-99."""
    from diffusesg_b200.loss.rainbow_loss import NodeAdjRainbowLoss
    from oracle import train_oracle as T
    g = torch.Generator().manual_seed(b * n)
    flags = torch.arange(n)[None, :] < torch.randint(2, n + 1, (b, 1), generator=g)
    pa, ta = torch.randn(b, ce, n, n, generator=g), torch.randn(b, ce, n, n, generator=g)
    pn, tn = torch.randn(b, n, cn, generator=g), torch.randn(b, n, cn, generator=g)
    w = torch.rand(b, generator=g) + 0.5
    loss = NodeAdjRainbowLoss(edge_loss_weight=1.0, node_loss_weight=0.5, objective="edm")
    ga = pa.clone().to(DEV).requires_grad_(True)
    gn = pn.clone().to(DEV).requires_grad_(True)
    la, ln = loss(ga, gn, ta.to(DEV), tn.to(DEV), None, None, None, None, None, flags.to(DEV), loss_weight=w.to(DEV), reduction=red)
    (la.sum() + 2.0 * ln.sum()).backward()
    ra, rn = pa.clone().requires_grad_(True), pn.clone().requires_grad_(True)
    oa, on = T.regression_loss(ra, rn, ta, tn, flags, w, 1.0, 0.5, red)
    (oa.sum() + 2.0 * on.sum()).backward()
    torch.testing.assert_close(la.detach().cpu(), oa.detach(), rtol=1e-5, atol=0)
    torch.testing.assert_close(ga.grad.cpu(), ra.grad, rtol=2e-5, atol=1e-9)
    torch.testing.assert_close(gn.grad.cpu(), rn.grad, rtol=2e-5, atol=1e-9)
    assert float(ga.grad.cpu()[~(flags[:, None, :, None] & flags[:, None, None, :]).expand_as(pa)].abs().sum()) == 0.0
