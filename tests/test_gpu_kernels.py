"""Kernel-level parity tests (B200 only): every C-ABI kernel family against a plain fp32 PyTorch
statement of the same op on the same seeded inputs.

Tolerances (stated per test): bf16-output kernels are compared to the fp32 result with
rel-L2 <= 4e-3 (one bf16 rounding is 2^-9 = 2e-3 per element); fp32-output kernels with
rel-L2 <= 1e-5 .. 1e-3 depending on accumulation length; integer / index / mask outputs are
bit-exact.
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

bf16 = torch.bfloat16
DEV = "cuda"


def K():
    from speech_ssl_compression_b200 import kernels

    return kernels


def rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(1234)
    K().set_dropout_offset(None)
    yield
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,Kd,a_mn,b_mn,bn", [
    (128, 128, 64, False, False, 128),
    (256, 256, 768, False, False, 256),
    (3000, 2304, 768, False, False, 0),      # fused QKV, cfg2
    (3000, 768, 3072, False, False, 0),      # fc2
    (1000, 2112, 80, False, False, 0),       # K = 80 (pre_extract_proj 20 ms), ragged N tile
    (333, 64, 40, False, False, 0),          # K = 40 (10 ms), one head
    (777, 1536, 768, False, False, 0),       # row-pruned fc1
    (640, 704, 768, False, False, 0),        # 11 heads
    (300, 768, 1536, False, True, 0),        # dgrad through a row-pruned fc1 (B MN-major)
    (256, 256, 256, True, False, 256),
    (256, 256, 256, True, True, 128),
    # CTA-pair (cta_group::2) 256 x 256 tiles, every operand layout, ragged M / N edges
    (256, 256, 64, False, False, -256),
    (3000, 2304, 768, False, False, -256),
    (1000, 704, 768, False, False, -256),
    (300, 768, 1536, False, True, -256),
    (512, 512, 256, True, False, -256),
    (776, 328, 192, True, True, -256),
])
def test_gemm_bf16_matches_fp32(M, N, Kd, a_mn, b_mn, bn):
    k = K()
    a = torch.randn(M, Kd, device=DEV).to(bf16)
    b = torch.randn(N, Kd, device=DEV).to(bf16)
    bias = torch.randn(N, device=DEV)
    ref = a.float() @ b.float().t() + bias
    out = torch.empty(M, N, device=DEV, dtype=bf16)
    k.gemm(a.t().contiguous() if a_mn else a, b.t().contiguous() if b_mn else b, out, a_mn=a_mn, b_mn=b_mn, bias=bias,
           block_n=bn)
    assert rel(out, ref) < 4e-3


@pytest.mark.parametrize("M,N,Kd,split,bn", [(768, 3072, 3000, 0, 0), (768, 768, 3000, 0, 0), (2304, 768, 24000, 0, 0),
                                             (512, 80, 3000, 3, 0), (768, 1536, 1454, 5, 0), (768, 3072, 3000, 0, 128),
                                             (704, 776, 1000, 2, -256)])
def test_gemm_wgrad_f32_accumulates_with_mask(M, N, Kd, split, bn):
    """dW += dY^T X with both operands MN-major, fp32 red.add epilogue, optional prune mask,
    split-K; checks the accumulate-into-existing-gradient semantics too."""
    k = K()
    dy = torch.randn(Kd, M, device=DEV).to(bf16)   # stored [K][M]
    x = torch.randn(Kd, N, device=DEV).to(bf16)    # stored [K][N]
    mask = torch.rand(M, N, device=DEV) > 0.5
    prev = torch.randn(M, N, device=DEV)
    ref = dy.float().t() @ x.float()
    out = prev.clone()
    k.gemm(dy, x, out, a_mn=True, b_mn=True, epilogue=k.EPI_F32, mask=mask, split_k=split, block_n=bn)
    want = prev + ref * mask
    assert rel(out, want) < 1e-4
    assert torch.equal(out[~mask], prev[~mask]), "masked-out gradient entries must be untouched"


def test_gemm_epilogues_gelu_res_dgelu_add():
    k = K()
    M, N, Kd = 1000, 3072, 768
    a = torch.randn(M, Kd, device=DEV).to(bf16)
    w = (torch.randn(N, Kd, device=DEV) * 0.03).to(bf16)
    bias = torch.randn(N, device=DEV) * 0.1
    acc = a.float() @ w.float().t() + bias
    # GELU: pre rounded to bf16 first, erf-GELU in fp32 (fairseq_code/gelu.py:35 under autocast)
    out = torch.empty(M, N, device=DEV, dtype=bf16)
    pre = torch.empty_like(out)
    k.gemm(a, w, out, epilogue=k.EPI_GELU, bias=bias, aux_out=pre)
    assert rel(pre, acc) < 4e-3
    assert rel(out, torch.nn.functional.gelu(pre.float())) < 4e-3
    # RES: acc + bias + residual
    res = torch.randn(M, N, device=DEV).to(bf16)
    k.gemm(a, w, out, epilogue=k.EPI_RES, bias=bias, aux_in=res)
    assert rel(out, acc + res.float()) < 4e-3
    # ADD (no bias)
    k.gemm(a, w, out, epilogue=k.EPI_ADD, aux_in=res)
    assert rel(out, acc - bias + res.float()) < 4e-3
    # DGELU: acc * gelu'(aux)
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    k.gemm(a, w, out, epilogue=k.EPI_DGELU, aux_in=pre)
    assert rel(out, (acc - bias) * x.grad) < 6e-3


@pytest.mark.parametrize("B,T,H,block_n", [(32, 750, 12, 0), (2, 300, 3, 0), (4, 192, 7, 256)])
def test_gemm_delta_epilogue_feeds_the_attention_backward(B, T, H, block_n):
    """MH_EPI_DELTA: the out_proj dgrad GEMM also emits delta = rowsum(dO * O) per (batch, head, query) -- same values
    as the stand-alone kernel inside mh_attn_bwd, and the attention backward gives the same gradients with it."""
    k = K()
    E = 64 * H
    M = B * T
    torch.manual_seed(3)
    dz = (torch.randn(M, 768, device=DEV) * 0.5).to(bf16)
    wo = (torch.randn(768, E, device=DEV) * 0.03).to(bf16)        # forward weight [out, in]: MN-major B of the dgrad
    qkv = torch.randn(M, 3 * E, device=DEV).to(bf16)
    lens_t = torch.tensor([T - 7 * (i % 3) for i in range(B)], device=DEV, dtype=torch.int32)
    ctx, lse, keep = k.attn_fwd(qkv, lens_t, B, T, H, p_drop=0.1, seed=5, site=1)
    dctx, dref = torch.empty(M, E, device=DEV, dtype=bf16), torch.empty(M, E, device=DEV, dtype=bf16)
    delta = torch.full((B, H, T), float("nan"), device=DEV)
    k.gemm(dz, wo, dctx, b_mn=True, epilogue=k.EPI_DELTA, aux_in=ctx, delta=delta, delta_T=T, block_n=block_n)
    k.gemm(dz, wo, dref, b_mn=True)
    assert torch.equal(dctx, dref)
    want = (dref.float() * ctx.float()).view(B, T, H, 64).sum(-1).permute(0, 2, 1)
    torch.testing.assert_close(delta, want, rtol=1e-4, atol=1e-4 * float(want.abs().max()))
    g_fused = k.attn_bwd(qkv, lens_t, ctx, dctx, lse, keep, B, T, H, p_drop=0.1, seed=5, site=1, delta=delta)
    g_plain = k.attn_bwd(qkv, lens_t, ctx, dctx, lse, keep, B, T, H, p_drop=0.1, seed=5, site=1)
    assert rel(g_fused, g_plain) < 2e-3  # (fp32 reduce-add order differs from launch to launch)
    # q / k / v bias gradients from the same call (mh_attn_bwd_bias): same dqkv; the k / v column sums come from the fp32
    # dK / dV accumulators inside the kernel (so they differ from sums of the bf16-rounded outputs by rounding noise), the q
    # sums from the bf16 conversion pass
    bg = torch.full((3 * E,), 1.0, device=DEV)
    g_b = k.attn_bwd(qkv, lens_t, ctx, dctx, lse, keep, B, T, H, p_drop=0.1, seed=5, site=1, delta=delta, bias_grad=bg)
    assert rel(g_b, g_plain) < 2e-3
    want_b = g_b.float().sum(0)
    scale = float(g_b.float().abs().sum(0).max())  # (column sums of random-sign gradients cancel: judge against sum |.|)
    torch.testing.assert_close(bg[:E] - 1.0, want_b[:E], rtol=1e-3, atol=1e-4 * scale + 1e-3)
    torch.testing.assert_close(bg[E:] - 1.0, want_b[E:], rtol=1e-3, atol=2e-3 * scale + 1e-3)
    assert rel(bg[E:] - 1.0, want_b[E:]) < 2e-2
    # the stand-alone finishing pass (mh_dq_finish_colsum: all 3E column sums from the stored tensor)
    dq_acc = torch.randn(M, E, device=DEV)
    g_c = g_b.clone()
    cs = torch.zeros(3 * E, device=DEV)
    k.dq_finish_colsum(dq_acc, g_c, cs)
    assert torch.equal(g_c[:, :E], dq_acc.to(bf16)) and torch.equal(g_c[:, E:], g_b[:, E:])
    torch.testing.assert_close(cs, g_c.float().sum(0), rtol=1e-3, atol=1e-4 * scale + 1e-3)
    with pytest.raises(RuntimeError):
        k.gemm(dz, wo.t().contiguous(), dctx, epilogue=k.EPI_DELTA, aux_in=ctx, delta=delta, delta_T=T)  # K-major B: not built


def test_gemm_dropout_epilogue_statistics_and_regeneration():
    """Dropout in the epilogue: keep-rate ~ 1-p, survivors scaled by 1/(1-p), the same
    (seed, site) regenerates the same mask, a different site gives a different one, and
    mh_dropout_apply reproduces the GEMM epilogue's mask element for element."""
    k = K()
    M, N, Kd, p = 2048, 768, 64, 0.1
    a = torch.ones(M, Kd, device=DEV).to(bf16)
    w = torch.full((N, Kd), 1.0 / Kd, device=DEV).to(bf16)
    zero = torch.zeros(M, N, device=DEV, dtype=bf16)
    o1, o2, o3 = (torch.empty(M, N, device=DEV, dtype=bf16) for _ in range(3))
    k.gemm(a, w, o1, epilogue=k.EPI_RES, aux_in=zero, p_drop=p, seed=77, site=5)
    k.gemm(a, w, o2, epilogue=k.EPI_RES, aux_in=zero, p_drop=p, seed=77, site=5)
    k.gemm(a, w, o3, epilogue=k.EPI_RES, aux_in=zero, p_drop=p, seed=77, site=6)
    assert torch.equal(o1, o2)
    assert not torch.equal(o1, o3)
    keep = (o1 != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 3e-3
    vals = o1[o1 != 0].float()
    assert torch.allclose(vals, torch.full_like(vals, 1 / (1 - p)), rtol=1e-2)
    ones = torch.ones(M, N, device=DEV, dtype=bf16)
    d = k.dropout_apply(ones, p, 77, 5)
    assert torch.equal(d != 0, o1 != 0)


def test_gemm_rejects_bad_arguments():
    k = K()
    a = torch.randn(64, 60, device=DEV).to(bf16)  # ld = 60 not a multiple of 8
    b = torch.randn(64, 60, device=DEV).to(bf16)
    out = torch.empty(64, 64, device=DEV, dtype=bf16)
    with pytest.raises(RuntimeError):
        k.gemm(a, b, out)
    with pytest.raises(ValueError):
        k.gemm(a.float(), b, out)


# -------------------------------------------------------------------------------------- attention
def _ref_attn(qkv, lens, B, T, H, causal):
    E = 64 * H
    x = qkv.view(B, T, 3, H, 64)
    q, kk, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    s = (q / 8.0) @ kk.transpose(-1, -2)
    ar = torch.arange(T, device=qkv.device)
    s = s.masked_fill((ar[None, :] >= lens[:, None])[:, None, None, :], float("-inf"))
    if causal:
        s = s.masked_fill(torch.ones(T, T, dtype=torch.bool, device=qkv.device).triu(1), float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, E)


@pytest.mark.parametrize("B,T,H,lens,causal", [
    (1, 128, 1, [128], False),
    (2, 300, 3, [300, 77], False),
    (4, 750, 12, [750, 712, 655, 601], False),     # cfg2 ragged batch
    (2, 791, 12, [101, 791], False),               # cfg1 extraction batch (unsorted lengths)
    (2, 1500, 7, [1500, 1203], False),             # 10 ms, 7 surviving heads
    (2, 300, 2, [300, 211], True),
])
def test_attention_fwd_bwd(B, T, H, lens, causal):
    k = K()
    E = 64 * H
    qkv = torch.randn(B * T, 3 * E, device=DEV).to(bf16)
    lens_t = torch.tensor(lens, device=DEV, dtype=torch.int32)
    out, lse, keep = k.attn_fwd(qkv, lens_t, B, T, H, causal=causal)
    qr = qkv.float().requires_grad_(True)
    ref = _ref_attn(qr, lens_t, B, T, H, causal)
    valid = (torch.arange(T, device=DEV)[None, :] < lens_t[:, None]).reshape(-1)
    assert rel(out[valid], ref[valid]) < 6e-3
    dout = torch.randn(B * T, E, device=DEV).to(bf16)
    dqkv = k.attn_bwd(qkv, lens_t, out, dout, lse, keep, B, T, H, causal=causal)
    ref.backward(dout.float())
    g = qr.grad
    for sl in (slice(0, E), slice(E, 2 * E), slice(2 * E, 3 * E)):
        assert rel(dqkv[:, sl], g[:, sl]) < 1.2e-2
    # padded keys receive exactly zero gradient
    if not bool(valid.all()):
        assert dqkv[~valid][:, E:].abs().max().item() == 0.0


def test_attention_dropout_is_regenerated_in_backward():
    """Finite-difference-free check: with dropout on, d(sum(out * dout))/dv computed by the
    kernel must equal P_drop^T dout, where P_drop is recovered from a forward with V = I."""
    k = K()
    B, T, H, p = 1, 128, 1, 0.25
    qkv = torch.randn(T, 192, device=DEV).to(bf16)
    lens_t = torch.tensor([T], device=DEV, dtype=torch.int32)
    o1, lse, keep = k.attn_fwd(qkv, lens_t, B, T, H, p_drop=p, seed=5, site=3)
    o2, _, _ = k.attn_fwd(qkv, lens_t, B, T, H, p_drop=p, seed=5, site=3)
    assert torch.equal(o1, o2)
    # recover the dropped probability matrix column block by column block with one-hot V
    P = torch.zeros(T, T, device=DEV)
    for c0 in range(0, T, 64):
        q2 = qkv.clone()
        q2[:, 128:] = 0
        q2[c0:c0 + 64, 128:] = torch.eye(64, device=DEV).to(bf16)
        oc, _, _ = k.attn_fwd(q2, lens_t, B, T, H, p_drop=p, seed=5, site=3)
        P[:, c0:c0 + 64] = oc.float()
    dropped = (P == 0).float().mean().item()
    assert abs(dropped - p) < 0.02
    dout = torch.randn(T, 64, device=DEV).to(bf16)
    dqkv = k.attn_bwd(qkv, lens_t, o1, dout, lse, keep, B, T, H, p_drop=p, seed=5, site=3)
    dv_ref = P.t() @ dout.float()
    assert rel(dqkv[:, 128:], dv_ref) < 1.5e-2


def test_attention_dropout_backward_all_gradients_multi_block():
    """Dropout on, several key / query blocks, ragged lengths, two heads (several work items per persistent CTA):
    the keep mask is recovered from forwards with one-hot V, then dQ, dK and dV must match autograd through
    softmax -> mask / (1 - p) -> PV with exactly that mask (the backward reads the keep bits the forward saved)."""
    k = K()
    B, T, H, p = 2, 300, 2, 0.2
    E = 64 * H
    lens = [300, 211]
    qkv = torch.randn(B * T, 3 * E, device=DEV).to(bf16)
    lens_t = torch.tensor(lens, device=DEV, dtype=torch.int32)
    o1, lse, keep = k.attn_fwd(qkv, lens_t, B, T, H, p_drop=p, seed=77, site=5)
    # P_drop[b, h, :, c0:c0+64] = output of a forward whose V rows c0.. are the identity (same seed -> same mask)
    Pd = torch.zeros(B, H, T, T, device=DEV)
    for c0 in range(0, T, 64):
        n = min(64, T - c0)
        q2 = qkv.clone().view(B, T, 3, H, 64)
        q2[:, :, 2] = 0
        q2[:, c0:c0 + n, 2] = torch.eye(64, device=DEV).to(bf16)[:n][None, :, None, :]
        oc, _, _ = k.attn_fwd(q2.view(B * T, 3 * E), lens_t, B, T, H, p_drop=p, seed=77, site=5)
        Pd[:, :, :, c0:c0 + n] = oc.float().view(B, T, H, 64).permute(0, 2, 1, 3)[..., :n]
    ar = torch.arange(T, device=DEV)
    vis = (ar[None, :] < lens_t[:, None])[:, None, None, :].expand(B, H, T, T)
    M = (Pd != 0) & vis
    rate = 1.0 - M[vis].float().mean().item()
    assert abs(rate - p) < 0.01
    qr = qkv.float().view(B, T, 3, H, 64).clone().requires_grad_(True)
    q, kk, v = (qr[:, :, i].transpose(1, 2) for i in range(3))
    s_ = (q @ kk.transpose(-1, -2)) * 0.125
    s_ = s_.masked_fill(~vis, float("-inf"))
    Pref = torch.softmax(s_, -1) * M / (1.0 - p)
    out_ref = (Pref @ v).transpose(1, 2).reshape(B * T, E)
    valid = (ar[None, :] < lens_t[:, None]).reshape(-1)
    assert rel(o1[valid], out_ref[valid]) < 8e-3
    dout = torch.randn(B * T, E, device=DEV).to(bf16)
    dout[~valid] = 0          # rows past an utterance's end carry no gradient in the model either
    dqkv = k.attn_bwd(qkv, lens_t, o1, dout, lse, keep, B, T, H, p_drop=p, seed=77, site=5)
    out_ref.backward(dout.float())
    g = qr.grad.reshape(B * T, 3 * E)
    for sl in (slice(0, E), slice(E, 2 * E), slice(2 * E, 3 * E)):
        assert rel(dqkv[valid][:, sl], g[valid][:, sl]) < 1.5e-2


def test_attention_full_size_properties():
    """cfg2 bench shape (B = 32, T = 750, 12 heads, dropout 0.1): too large for the materialised reference, so the
    size-independent properties -- bit-identical replays, keep rate, and linearity of the backward in dO."""
    k = K()
    B, T, H, p = 32, 750, 12, 0.1
    E = 64 * H
    torch.manual_seed(3)
    qkv = torch.randn(B * T, 3 * E, device=DEV).to(bf16)
    lens = sorted([T - int(x) for x in torch.randint(0, T // 5, (B,))], reverse=True)
    lens_t = torch.tensor(lens, device=DEV, dtype=torch.int32)
    o1, lse1, keep1 = k.attn_fwd(qkv, lens_t, B, T, H, p_drop=p, seed=9, site=2)
    o2, lse2, keep2 = k.attn_fwd(qkv, lens_t, B, T, H, p_drop=p, seed=9, site=2)
    assert torch.equal(o1, o2) and torch.equal(lse1, lse2)
    o0, lse0, _ = k.attn_fwd(qkv, lens_t, B, T, H)
    valid = (torch.arange(T, device=DEV)[None, :] < lens_t[:, None]).reshape(-1)
    assert torch.isfinite(o1[valid].float()).all()
    torch.testing.assert_close(lse1[:, :, :lens[-1]], lse0[:, :, :lens[-1]], rtol=0, atol=2e-3)  # lse is dropout-free
    assert rel(o1[valid], o0[valid]) < 0.5    # dropout noise on top of the same expectation
    # keep bits: one bit per (row, visible key); population count / visible scores = 1 - p
    words = keep1.view(torch.int32).view(B, H, -1, T)          # [B, H, 4 * ceil(T / 128), T] query-minor
    nw = (lens[0] + 31) // 32
    w0 = words[0, :, :nw - 1, :lens[0]]                         # full words of the longest utterance
    bits = sum(((w0 >> i) & 1).sum().item() for i in range(32))
    assert abs(bits / (w0.numel() * 32) - (1 - p)) < 2e-3
    d1 = torch.randn(B * T, E, device=DEV).to(bf16)
    d2 = torch.randn(B * T, E, device=DEV).to(bf16)
    g1 = k.attn_bwd(qkv, lens_t, o1, d1, lse1, keep1, B, T, H, p_drop=p, seed=9, site=2).float()
    g2 = k.attn_bwd(qkv, lens_t, o1, d2, lse1, keep1, B, T, H, p_drop=p, seed=9, site=2).float()
    g12 = k.attn_bwd(qkv, lens_t, o1, (d1.float() + d2.float()).to(bf16), lse1, keep1, B, T, H, p_drop=p, seed=9, site=2).float()
    assert rel(g12[valid], (g1 + g2)[valid]) < 1.5e-2
    g1b = k.attn_bwd(qkv, lens_t, o1, d1, lse1, keep1, B, T, H, p_drop=p, seed=9, site=2).float()
    assert rel(g1b[:, E:], g1[:, E:]) == 0.0 or rel(g1b[:, E:], g1[:, E:]) < 1e-6   # dK / dV: no atomics
    assert rel(g1b[:, :E], g1[:, :E]) < 1e-3                                        # dQ: fp32 reduce-add order


# ------------------------------------------------------------------------------------- layer norm
@pytest.mark.parametrize("rows,cols", [(3000, 768), (257, 768), (64, 512), (24000, 768), (5000, 80), (9001, 1000), (1, 768)])
def test_layernorm_fwd_bwd(rows, cols):
    k = K()
    x = (torch.randn(rows, cols, device=DEV) * 2 + 0.5).to(bf16)
    g = torch.randn(cols, device=DEV)
    b = torch.randn(cols, device=DEV)
    y, mean, rstd = k.layernorm_fwd(x, g, b, 1e-5)
    xr = x.float().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (cols,), gr, br, 1e-5)
    assert rel(y, ref) < 4e-3
    assert rel(mean, xr.mean(-1)) < 1e-5
    dy = torch.randn(rows, cols, device=DEV).to(bf16)
    ref.backward(dy.float())
    dg, db = torch.zeros(cols, device=DEV), torch.zeros(cols, device=DEV)
    dx, _ = k.layernorm_bwd(dy, x, g, mean, rstd, dg, db)
    assert rel(dx, xr.grad) < 5e-3
    assert rel(dg, gr.grad) < 1e-3
    assert rel(db, br.grad) < 1e-3


@pytest.mark.parametrize("rows,cols", [(512, 768), (1237, 544), (300, 160), (2500, 1024)])
def test_layernorm_output_dropout_and_masked_input_gradient(rows, cols):
    # (544 / 160 columns: partly filled last 256-column chunk group, i.e. idle lanes and fewer dropout words in the last
    #  warp of a row; 1237 / 300 / 2500 rows: row slots of a block run out at different trips, 4-row keep-word batches)
    k = K()
    p = 0.1
    x = torch.randn(rows, cols, device=DEV).to(bf16)
    g, b = torch.ones(cols, device=DEV), torch.zeros(cols, device=DEV)
    y0, mean, rstd = k.layernorm_fwd(x, g, b, 1e-5)
    y1, _, _ = k.layernorm_fwd(x, g, b, 1e-5, p_drop=p, seed=9, site=4)
    keep = y1 != 0
    assert abs(keep.float().mean().item() - (1 - p)) < 5e-3
    assert rel(y1[keep], y0[keep].float() / (1 - p)) < 6e-3
    # backward with p_in: gradient flows only through kept elements, scaled by 1/(1-p)
    dy = torch.randn(rows, cols, device=DEV).to(bf16)
    dg, db = torch.zeros(cols, device=DEV), torch.zeros(cols, device=DEV)
    dx, _ = k.layernorm_bwd(dy, x, g, mean, rstd, dg, db, p_in=p, seed_in=9, site_in=4)
    dy_eff = (dy.float() * keep / (1 - p)).to(bf16)
    dg2, db2 = torch.zeros(cols, device=DEV), torch.zeros(cols, device=DEV)
    dx2, _ = k.layernorm_bwd(dy_eff, x, g, mean, rstd, dg2, db2)
    assert rel(dx, dx2) < 8e-3
    # p_out: second output = dx * keep-mask(site_out) / (1 - p)
    dxa, dxd = k.layernorm_bwd(dy, x, g, mean, rstd, dg, db, want_drop=True, p_out=p, seed_out=9, site_out=11)
    m = k.dropout_apply(torch.ones_like(x), p, 9, 11) != 0
    assert rel(dxd, dxa.float() * m / (1 - p)) < 6e-3
    # fused bias gradient: column sums of the second output (of dx when there is none), accumulated in place
    for want_drop in (True, False):
        cs = torch.full((cols,), 3.0, device=DEV)
        a1, d1 = k.layernorm_bwd(dy, x, g, mean, rstd, dg, db, want_drop=want_drop, p_out=p if want_drop else 0.0,
                                 seed_out=9, site_out=11, colsum_out=cs)
        src = d1 if want_drop else a1
        torch.testing.assert_close(cs - 3.0, src.float().sum(0), rtol=2e-2, atol=2e-2 * float(src.float().sum(0).abs().max()))
        assert torch.equal(a1, dxa)


def test_colsum_and_casts():
    k = K()
    x = torch.randn(3000, 2304, device=DEV).to(bf16)
    out = torch.ones(768, device=DEV)
    k.colsum_add(x[:, 768:1536], out)
    assert rel(out, 1 + x[:, 768:1536].float().sum(0)) < 1e-4
    # narrow + tall matrices take the row-per-warp kernel (4 rows in flight, ragged tail, strided view, odd widths)
    for rows, cols, ld in [(24000, 768, 768), (4099, 768, 2304), (9001, 1000, 1000), (5000, 80, 80)]:
        big = torch.randn(rows, ld, device=DEV).to(bf16)
        view = big[:, ld - cols:] if ld != cols else big
        acc = torch.full((cols,), 2.0, device=DEV)
        k.colsum_add(view, acc)
        assert rel(acc, 2 + view.float().sum(0)) < 1e-4, (rows, cols, ld)
    f = torch.randn(1000, 80, device=DEV)
    assert torch.equal(k.to_bf16(f), f.to(bf16))
    assert torch.equal(k.to_f32(f.to(bf16)), f.to(bf16).float())


# ------------------------------------------------------------------------------- prep / selection
def test_weight_prep_masks_and_transposes():
    k = K()
    w = torch.randn(704, 768, device=DEV)
    m = torch.rand(704, 768, device=DEV) > 0.5
    dst = torch.empty(2112, 768, device=DEV, dtype=bf16)
    dst_t = torch.empty(768, 704, device=DEV, dtype=bf16)
    k.weight_prep(w, m, dst[704:1408], dst_t)
    want = w.masked_fill(~m, 0).to(bf16)   # prune.py:83-84 apply_mask, then the autocast cast
    assert torch.equal(dst[704:1408], want)
    assert torch.equal(dst_t, want.t())
    k.weight_prep(w, None, dst[:704])
    assert torch.equal(dst[:704], w.to(bf16))
    bsrc, bm = torch.randn(704, device=DEV), torch.rand(704, device=DEV) > 0.3
    bdst = torch.empty(704, device=DEV)
    k.bias_prep(bsrc, bm, bdst)
    assert torch.equal(bdst, bsrc.masked_fill(~bm, 0))


def test_select_gather_scatter_are_bit_exact():
    """model.py:147-150: hidden[masked_indices] / label[masked_indices] in row-major (b, t) order."""
    k = K()
    rows, cols = 3000, 768
    sel = (torch.rand(rows, device=DEV) > 0.52)
    idx, count = k.select_rows(sel.to(torch.uint8))
    n = int(count.item())
    want = torch.nonzero(sel).flatten().to(torch.int32)
    assert n == want.numel()
    assert torch.equal(idx[:n], want)
    assert (idx[n:] == -1).all()
    x = torch.randn(rows, cols, device=DEV).to(bf16)
    assert torch.equal(k.gather_rows(x, idx, n), x[sel])
    label = torch.randint(0, 512, (rows,), device=DEV)
    assert torch.equal(k.gather_labels(label, idx, n), label[sel])
    # static-row variant: padded tail rows come back as zeros / label -100
    full = k.gather_rows(x, idx, rows)
    assert torch.equal(full[:n], x[sel]) and full[n:].abs().max().item() == 0
    assert (k.gather_labels(label, idx, rows)[n:] == -100).all()
    g = torch.randn(n, cols, device=DEV).to(bf16)
    dx = torch.zeros(rows, cols, device=DEV, dtype=bf16)
    k.scatter_rows_add_(g, idx, n, dx)
    ref = torch.zeros_like(dx)
    ref[sel] = g
    assert torch.equal(dx, ref)
    # empty and full selections
    for s in (torch.zeros(rows, device=DEV, dtype=torch.uint8), torch.ones(rows, device=DEV, dtype=torch.uint8)):
        i2, c2 = k.select_rows(s)
        assert int(c2.item()) == int(s.sum().item())


def test_mask_rows_and_zero_rows():
    k = K()
    x = torch.randn(1500, 80, device=DEV)
    z = (torch.rand(1500, device=DEV) > 0.5).to(torch.uint8)
    got = k.mask_rows_to_bf16(x, z)
    assert torch.equal(got, x.masked_fill(z.bool()[:, None], 0).to(bf16))
    assert torch.equal(k.mask_rows_to_bf16(x, None), x.to(bf16))
    y = torch.randn(1500, 768, device=DEV).to(bf16)
    want = y.masked_fill(z.bool()[:, None], 0)
    k.zero_rows_(y, z)
    assert torch.equal(y, want)


# --------------------------------------------------------------------------------------- criteria
def test_cross_entropy_fwd_bwd():
    """CrossEntropyLoss(ignore_index=-100, reduction='mean'), pretrain_expert.py:25,116-117."""
    from speech_ssl_compression_b200 import ops

    n, c = 1454, 512
    logits = (torch.randn(n, c, device=DEV) * 3).to(bf16)
    labels = torch.randint(0, c, (n,), device=DEV)
    labels[::17] = -100
    lg = logits.clone().requires_grad_(True)
    loss = ops.cross_entropy(lg, labels)
    loss.backward()
    lr = logits.float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, labels, ignore_index=-100)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
    assert rel(lg.grad, lr.grad) < 6e-3
    assert lg.grad[::17].abs().max().item() == 0
    # n_valid limits the rows that count (static-shape CUDA-graph path)
    nv = torch.tensor([1000], device=DEV, dtype=torch.int32)
    l2 = ops.cross_entropy(logits, labels, n_valid=nv)
    r2 = torch.nn.functional.cross_entropy(logits[:1000].float(), labels[:1000], ignore_index=-100)
    assert abs(l2.item() - r2.item()) < 1e-4 * abs(r2.item())


@pytest.mark.parametrize("T,alpha", [(1.0, 1.0), (1.0, 0.5), (2.0, 0.3)])
def test_kd_loss_fwd_bwd(T, alpha):
    """distillation/pretrain_expert.py:83-92."""
    from speech_ssl_compression_b200 import ops

    n, c = 3000, 512
    s = (torch.randn(n, c, device=DEV) * 2).to(bf16)
    t = (torch.randn(n, c, device=DEV) * 2).to(bf16)
    labels = torch.randint(0, c, (n,), device=DEV)
    # rows labelled -100 are the padding slots of a statically shaped row list: they count in
    # neither term (label_m / label_u of the reference never hold -100: model.py:147 drops padded frames)
    labels[-300:] = -100
    sg = s.clone().requires_grad_(True)
    total, terms = ops.kd_loss(sg, t, labels, T, alpha)
    total.backward()
    sr = s.float().requires_grad_(True)
    F = torch.nn.functional
    ok = labels >= 0
    hard = F.cross_entropy(sr[ok], labels[ok])
    soft = torch.nn.KLDivLoss(reduction="batchmean")(F.log_softmax(sr[ok] / T, dim=1), F.softmax(t.float()[ok] / T, dim=1))
    tce = F.cross_entropy(t.float()[ok], labels[ok])
    ref = (1 - alpha) * hard + alpha * soft
    ref.backward()
    assert abs(total.item() - ref.item()) < 2e-4 * max(1.0, abs(ref.item()))
    assert abs(terms[1].item() - hard.item()) < 2e-4 * hard.item()
    assert abs(terms[2].item() - soft.item()) < 2e-4 * max(1.0, soft.item())
    assert abs(terms[3].item() - tce.item()) < 2e-4 * tce.item()
    assert rel(sg.grad, sr.grad) < 8e-3
    assert sg.grad[-300:].abs().max().item() == 0


def test_l1_cosine_loss_fwd_bwd():
    from speech_ssl_compression_b200 import ops

    rows, cols = 3000, 768
    p = torch.randn(rows, cols, device=DEV).to(bf16)
    t = (p.float() * 0.5 + torch.randn(rows, cols, device=DEV)).to(bf16)
    pg = p.clone().requires_grad_(True)
    loss = ops.l1_cosine_loss(pg, t, 0.7)
    loss.backward()
    pr = p.float().requires_grad_(True)
    F = torch.nn.functional
    ref = F.l1_loss(pr, t.float()) + 0.7 * (-F.logsigmoid(F.cosine_similarity(pr, t.float(), dim=-1))).mean()
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-4 * abs(ref.item())
    assert rel(pg.grad, pr.grad) < 8e-3


# ---------------------------------------------------------------------------------- pruning objects
def test_global_threshold_masks_match_topk_outside_ties():
    """prune.py:553-573 (L1Unstructured.compute_mask via global_unstructured :1049-1171):
    k = round(amount * N) smallest |w| are cleared.  Exact outside the tie set, equal count inside."""
    k = K()
    torch.manual_seed(5)
    shapes = [(768, 768), (3072, 768), (768,), (768, 3072), (3072,)]
    ws = [torch.randn(*s, device=DEV) * 0.02 for s in shapes]
    ws[2].zero_()  # zero biases, like random init
    # plant exact ties at the threshold region
    flat = torch.cat([w.flatten() for w in ws])
    n = flat.numel()
    kk = int(round(0.5 * n))
    res = k.abs_kth_smallest(ws, kk)
    thr_bits = int(res[0].item())
    thr = np.array([thr_bits], dtype=np.uint32).view(np.float32)[0]
    ref_sorted = flat.abs().sort().values
    assert thr == ref_sorted[kk - 1].item()
    assert int(res[1].item()) == int((flat.abs() < float(thr)).sum().item())
    assert int(res[2].item()) == int((flat.abs() == float(thr)).sum().item())
    masks = [torch.ones(w.shape, device=DEV, dtype=torch.uint8) for w in ws]
    k.apply_threshold_masks(ws, masks, res, kk)
    got = torch.cat([m.flatten() for m in masks]).bool()
    assert int((~got).sum().item()) == kk
    assert (~got[flat.abs() < float(thr)]).all()
    assert got[flat.abs() > float(thr)].all()
    # ties resolved by lowest flat index
    ties = torch.nonzero(flat.abs() == float(thr)).flatten()
    n_tie_pruned = kk - int(res[1].item())
    assert (~got[ties[:n_tie_pruned]]).all() and got[ties[n_tie_pruned:]].all()


def test_row_and_col_abs_sums_fp64():
    k = K()
    w = torch.randn(3072, 768, device=DEV)
    assert torch.allclose(k.row_abs_sums(w), w.double().abs().sum(1), rtol=1e-12)
    assert torch.allclose(k.col_abs_sums(w), w.double().abs().sum(0), rtol=1e-12)
    v = w[:704]  # a row slice (heads 0..10)
    assert torch.allclose(k.row_abs_sums(v), v.double().abs().sum(1), rtol=1e-12)


# ---------------------------------------------------------------------------------------- optimizer
def test_fused_adam_matches_torch_adam_with_clipping():
    """runner.py:411-427: grads /= n; clip_grad_norm_(10); Adam.step; zero_grad."""
    k = K()
    n = 1_000_000 + 32
    p0 = torch.randn(n, device=DEV)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    step = torch.zeros(1, device=DEV, dtype=torch.int64)
    sumsq = torch.zeros(1, device=DEV)
    for it in range(3):
        g = torch.randn(n, device=DEV) * (0.05 if it else 1.0)  # first step clips (norm 1000 > 10)
        ref_p.grad = (g / 2.0).clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 10.0)
        opt.step()
        grad = g.clone()
        k.counter_add(step, 1)
        sumsq.zero_()
        k.sumsq_add(grad, sumsq)
        k.adam_step(p, grad, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=0.5, max_norm=10.0,
                    sumsq=sumsq, zero_grad=True)
        assert grad.abs().max().item() == 0
        assert rel(p, ref_p) < 1e-6
    assert (p - ref_p).abs().max().item() < 2e-6


# ------------------------------------------------------------------------------- positional conv
@pytest.mark.parametrize("B,T", [(1, 128), (2, 300), (4, 750), (2, 791), (1, 1500)])
def test_positional_conv_fwd_bwd_vs_torch(B, T):
    """module.py:175-188,229-231: x + GELU(SamePad(weight-normed grouped Conv1d(x))) and its three gradients."""
    F = torch.nn.functional
    k = K()
    C, groups, taps = 768, 16, 128
    v = torch.randn(C, C // groups, taps, device=DEV) * 0.02
    g = torch.rand(1, 1, taps, device=DEV) + 0.5
    bias = torch.randn(C, device=DEV) * 0.1
    x = torch.randn(B * T, C, device=DEV).to(bf16)
    w_fwd, w_bwd, norm = k.posconv_weight_prep(v, g.reshape(-1))
    y, z = k.posconv_fwd(x, w_fwd, bias, B, T)
    vr, gr = v.clone().requires_grad_(True), g.clone().requires_grad_(True)
    w = torch._weight_norm(vr, gr, 2)
    xr = x.float().view(B, T, C).requires_grad_(True)
    zr = F.conv1d(xr.transpose(1, 2), w, bias, padding=taps // 2, groups=groups)[:, :, :-1].transpose(1, 2)
    yr = xr + F.gelu(zr)
    assert rel(z, zr.reshape(B * T, C)) < 4e-3 and rel(y, yr.reshape(B * T, C)) < 4e-3
    dy = torch.randn(B * T, C, device=DEV).to(bf16)
    yr.backward(dy.float().view(B, T, C))
    dz = k.gelu_bwd_mul(dy, z)
    dx = k.posconv_dgrad(dz, w_bwd, dy, B, T)
    assert rel(dx, xr.grad.reshape(B * T, C)) < 5e-3
    dw = torch.zeros(C, C // groups, taps, device=DEV)
    k.posconv_wgrad(dz, x, dw, B, T)
    dv, dg = torch.zeros_like(v), torch.zeros(taps, device=DEV)
    k.posconv_weight_bwd(dw, v, g.reshape(-1), norm, dv, dg)
    assert rel(dv, vr.grad) < 6e-3 and rel(dg, gr.grad.reshape(-1)) < 6e-3
    db = torch.zeros(C, device=DEV)
    k.colsum_add(dz, db)
    zr2 = zr.detach().requires_grad_(True)
    F.gelu(zr2).backward(dy.float().view(B, T, C))
    assert rel(db, zr2.grad.sum((0, 1))) < 6e-3
