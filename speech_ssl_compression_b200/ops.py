"""Autograd glue between the reference-shaped module tree and the sm_100a kernels.

One ``torch.autograd.Function`` per fused block.  Activations are bf16 ``[B*T, C]`` matrices
(one row per frame, batch-major); master weights / gradients are fp32.  Weight gradients are
accumulated **in place** into ``param.grad`` by the wgrad GEMM epilogue (red.add), so gradient
accumulation over micro-batches and split-K are the same code path, and a data-parallel
wrapper can all-reduce a layer's gradients as soon as that layer's backward returns.

Dropout is regenerated in the backward from (seed, site, element index); nothing but the
activations listed in DESIGN.md is saved.
"""
import torch

from . import kernels as K

bf16 = torch.bfloat16

# ---------------------------------------------------------------------------------------------
# weight operand cache: fp32 master (+ prune mask) -> bf16, refreshed when the master changes
# ---------------------------------------------------------------------------------------------
_EPOCH = [0]


def bump_weight_epoch():
    """Call after any out-of-band parameter update (raw-pointer optimizer kernels, CUDA graph
    replays): forces the bf16 operands to be rebuilt on next use."""
    _EPOCH[0] += 1


def param_and_mask(mod, name):
    """(fp32 Parameter that trains, bool mask or None) for ``mod.<name>`` -- understands the
    ``<name>_orig`` / ``<name>_mask`` re-parametrisation of pytorch_code.prune."""
    if name + "_orig" in mod._parameters:
        return mod._parameters[name + "_orig"], mod._buffers[name + "_mask"]
    return mod._parameters[name], None


def _sig(tensors):
    return tuple((t.data_ptr(), t._version, tuple(t.shape)) if t is not None else None for t in tensors)


def packed_operands(owner, key, linears):
    """bf16 weight ``[sum N_i, K]`` (rows of the given linears stacked) and fp32 bias
    ``[sum N_i]`` with prune masks folded in.  Cached on ``owner``."""
    cache = owner.__dict__.setdefault("_mh_operands", {})
    parts = []
    for lin in linears:
        w, wm = param_and_mask(lin, "weight")
        b, bm = param_and_mask(lin, "bias")
        parts.append((w, wm, b, bm))
    shadow = _shadow_operands(parts)
    if shadow is not None:
        return shadow
    trainable = any(t is not None and t.requires_grad for p in parts for t in p)
    sig = (_EPOCH[0] if trainable else -1, _sig([t for p in parts for t in p]))
    hit = cache.get(key)
    if hit is not None and hit[0] == sig:
        return hit[1], hit[2]
    n_total = sum(p[0].shape[0] for p in parts)
    kdim = parts[0][0].shape[1]
    dev = parts[0][0].device
    if hit is not None and hit[1].shape == (n_total, kdim):
        wbuf, bbuf = hit[1], hit[2]
    else:
        wbuf = torch.empty(n_total, kdim, device=dev, dtype=bf16)
        bbuf = torch.empty(n_total, device=dev, dtype=torch.float32)
    with torch.no_grad():
        r = 0
        for w, wm, b, bm in parts:
            n = w.shape[0]
            K.weight_prep(w.detach(), wm, wbuf[r:r + n])
            K.bias_prep(b.detach(), bm, bbuf[r:r + n])
            r += n
    cache[key] = (sig, wbuf, bbuf)
    return wbuf, bbuf


def grad_hook(module_or_params, device):
    """A fresh zero-size leaf that stands in for a module's parameters in the autograd graph.

    The fused functions write parameter gradients in place (wgrad epilogues reduce-add into ``param.grad``), so
    the parameters themselves never need to be autograd inputs; passing them would also tie the graph to their
    long-lived AccumulateGrad nodes, whose creation stream breaks CUDA-graph capture once any other tensor
    (e.g. the ``weight = weight_orig * mask`` buffer of pytorch_code.prune) has kept them alive.  The hook is
    only created when some parameter is trainable, so frozen models (the distillation teacher) build no graph."""
    params = module_or_params.parameters() if hasattr(module_or_params, "parameters") else module_or_params
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return torch.empty(0, device=device, requires_grad=True)
    return None


def _shadow_operands(parts):
    """Fast path for trainable, unmasked weights that live in a FlatBuffers with a bf16 shadow kept current by the
    fused optimizer (mh_adam_step writes it in its own pass): the GEMM operand is a VIEW of the shadow -- no per-step
    weight prep.  Needs the tensors of `parts` back to back in the flat buffer (q, k, v are: trainer.trainable_params)."""
    flat0 = getattr(parts[0][0], "_mh_flat", None)
    if flat0 is None or flat0[0].flat_bf16 is None:
        return None
    fb, off = flat0
    if getattr(fb, "_shadow_epoch", None) != _EPOCH[0]:
        # parameters were changed outside the fused optimizer (load_state_dict, broadcast, surgery): re-sync
        fb.sync_shadow()
        fb._shadow_epoch = _EPOCH[0]
    kdim = parts[0][0].shape[1]
    w_off, b_off = off, None
    masked = False
    for w, wm, b, bm in parts:
        fw, fbias = getattr(w, "_mh_flat", None), getattr(b, "_mh_flat", None) if b is not None else None
        if b is None or fw is None or fbias is None or fw[0] is not fb or fbias[0] is not fb:
            return None
        if wm is not None or bm is not None:
            # prune masks: the shadow holds param * mask as long as the flat mask buffer mirrors these very mask tensors
            if not fb.owners:
                return None
            if (wm is not None and not fb.masks_current(w, wm)) or (bm is not None and not fb.masks_current(b, bm)):
                fb.sync_shadow()  # masks were replaced / edited since the last sync (prune event)
                if (wm is not None and not fb.masks_current(w, wm)) or (bm is not None and not fb.masks_current(b, bm)):
                    return None
            masked = True
        if fw[1] != w_off or w.shape[1] != kdim or w.data_ptr() != fb.flat_param.data_ptr() + 4 * fw[1]:
            return None  # not adjacent, or the Parameter was re-pointed (surgery) since the buffers were built
        if b_off is None:
            b_off = fbias[1]
        if fbias[1] != b_off or b.data_ptr() != fb.flat_param.data_ptr() + 4 * fbias[1]:
            return None
        w_off += w.numel()
        b_off += b.numel()
    n_total = sum(p[0].shape[0] for p in parts)
    wv = torch.as_strided(fb.flat_bf16, (n_total, kdim), (kdim, 1), off)
    bv = torch.as_strided(fb.flat_eff if (masked or fb.flat_eff is not None) else fb.flat_param, (n_total,), (1,),
                          parts[0][2]._mh_flat[1])
    return wv, bv


def _grad_of(p):
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


def _flat_masked(p):
    """True when ``p`` lives in a FlatBuffers whose optimizer applies the prune masks (flat_mask built)."""
    f = getattr(p, "_mh_flat", None)
    return f is not None and f[0].flat_mask is not None


def _bias_grad_target(lin):
    """``lin.bias.grad`` when its column sums may be accumulated by another kernel (the LayerNorm backward that produces
    the linear layer's output gradient): trainable, and no prune mask to apply here."""
    b, bm = param_and_mask(lin, "bias")
    if b is None or not b.requires_grad or (bm is not None and not _flat_masked(b)):
        return None
    return _grad_of(b)


def _wgrad(dy, x, lin, col0=0, ncols=None, bias_done=False):
    """lin.weight.grad += dy[:, col0:col0+n]^T @ x  (masked);  lin.bias.grad += column sums (unless ``bias_done``)."""
    w, wm = param_and_mask(lin, "weight")
    b, bm = param_and_mask(lin, "bias")
    n = w.shape[0] if ncols is None else ncols
    dyv = dy[:, col0:col0 + n] if (col0 or n != dy.shape[1]) else dy
    if _flat_masked(w):
        wm = bm = None  # the masked optimizer kernels drop the gradients of pruned elements (mh_adam_step_masked)
    if w.requires_grad:
        K.gemm(dyv, x, _grad_of(w), a_mn=True, b_mn=True, epilogue=K.EPI_F32, mask=wm)
    if b is not None and b.requires_grad and not bias_done:
        g = _grad_of(b)
        if bm is None:
            K.colsum_add(dyv, g)
        else:
            tmp = torch.zeros_like(g)
            K.colsum_add(dyv, tmp)
            g.add_(tmp * bm)


def _qkv_bias_grad_target(mha, E):
    """The stacked [3E] q / k / v bias gradient when the three sit back to back in the flat buffer and need no mask
    here: the attention backward's finishing pass accumulates it (``K.attn_bwd(bias_grad=...)``)."""
    gb = []
    for lin in (mha.q_proj, mha.k_proj, mha.v_proj):
        b, bm = param_and_mask(lin, "bias")
        if b is None or not b.requires_grad or b.grad is None or (bm is not None and not _flat_masked(b)) or b.numel() != E:
            return None
        gb.append(b.grad)
    if gb[1].data_ptr() != gb[0].data_ptr() + 4 * E or gb[2].data_ptr() != gb[1].data_ptr() + 4 * E:
        return None
    return torch.as_strided(gb[0], (3 * E,), (1,))


def _wgrad_qkv(dqkv, x, mha, E, bias_done=False):
    """Weight / bias gradients of the q, k, v projections.  When their gradients sit back to back in the flat
    buffer (trainer.trainable_params) and carry no prune masks: ONE [3E, C] wgrad GEMM and one column sum."""
    lins = (mha.q_proj, mha.k_proj, mha.v_proj)
    pw = [param_and_mask(l, "weight") for l in lins]
    pb = [param_and_mask(l, "bias") for l in lins]
    if all(_flat_masked(w) for w, _ in pw):
        pw = [(w, None) for w, _ in pw]
        pb = [(b, None) for b, _ in pb]
    fused = all(w.requires_grad and m is None and w.grad is not None for w, m in pw)
    if fused:
        g = [w.grad for w, _ in pw]
        n = g[0].numel()
        fused = (g[0].is_contiguous() and g[1].data_ptr() == g[0].data_ptr() + 4 * n
                 and g[2].data_ptr() == g[1].data_ptr() + 4 * n and g[0].shape == g[1].shape == g[2].shape)
    if not fused:
        for i, lin in enumerate(lins):
            _wgrad(dqkv, x, lin, col0=i * E, ncols=E, bias_done=bias_done)
        return
    C = g[0].shape[1]
    K.gemm(dqkv, x, torch.as_strided(g[0], (3 * E, C), (C, 1)), a_mn=True, b_mn=True, epilogue=K.EPI_F32)
    if bias_done:
        return
    gb = [b.grad if (b is not None and b.requires_grad and m is None) else None for b, m in pb]
    if all(t is not None for t in gb) and gb[1].data_ptr() == gb[0].data_ptr() + 4 * E \
            and gb[2].data_ptr() == gb[1].data_ptr() + 4 * E:
        K.colsum_add(dqkv, torch.as_strided(gb[0], (3 * E,), (1,)))
    else:
        for i, lin in enumerate(lins):
            b, bm = pb[i]
            if b is not None and b.requires_grad:
                gi = _grad_of(b)
                dyv = dqkv[:, i * E:(i + 1) * E]
                if bm is None:
                    K.colsum_add(dyv, gi)
                else:
                    tmp = torch.zeros_like(gi)
                    K.colsum_add(dyv, tmp)
                    gi.add_(tmp * bm)


# ---------------------------------------------------------------------------------------------
# plain linear (pre_extract_proj, final_proj, prediction heads)
# ---------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lin, hook):
        w, bias = packed_operands(lin, "self", [lin])
        out = torch.empty(x.shape[0], w.shape[0], device=x.device, dtype=bf16)
        K.gemm(x, w, out, bias=bias)
        ctx.lin = lin
        ctx.save_for_backward(x, w)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        _wgrad(dy, x, ctx.lin)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            K.gemm(dy, w, dx, b_mn=True)
        return dx, None, None


def linear(x, lin):
    """y = x @ W^T + b on the tcgen05 GEMM.  x: bf16 [rows, in_features]."""
    return LinearFn.apply(x, lin, grad_hook(lin, x.device))


# ---------------------------------------------------------------------------------------------
# LayerNorm (+ output dropout) -- encoder-level LN (module.py:232-236)
# ---------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ln, p_drop, seed, site, hook):
        y, mean, rstd = K.layernorm_fwd(x, ln.weight.detach(), ln.bias.detach(), ln.eps, p_drop=p_drop, seed=seed, site=site)
        ctx.ln, ctx.drop = ln, (p_drop, seed, site)
        ctx.save_for_backward(x, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        ln = ctx.ln
        p, seed, site = ctx.drop
        dx, _ = K.layernorm_bwd(dy.contiguous(), x, ln.weight.detach(), mean, rstd, _grad_of(ln.weight), _grad_of(ln.bias),
                                p_in=p, seed_in=seed, site_in=site)
        return dx, None, None, None, None, None


def layer_norm(x, ln, p_drop=0.0, seed=0, site=0):
    return LayerNormFn.apply(x, ln, p_drop, seed, site, grad_hook(ln, x.device))


# ---------------------------------------------------------------------------------------------
# positional convolution (module.py:175-188, 229-231): x + GELU(SamePad(Conv1d_wn(x)))
# ---------------------------------------------------------------------------------------------
def posconv_supported(conv):
    return (isinstance(conv, torch.nn.Conv1d) and conv.kernel_size == (128,) and conv.padding == (64,)
            and conv.in_channels == conv.out_channels and conv.in_channels // conv.groups == 48
            and hasattr(conv, "weight_v") and hasattr(conv, "weight_g"))


def _posconv_operands(conv):
    v, g = conv.weight_v, conv.weight_g
    trainable = v.requires_grad or g.requires_grad
    sig = (_EPOCH[0] if trainable else -1, _sig([v, g]))
    hit = conv.__dict__.get("_mh_posconv")
    if hit is not None and hit[0] == sig:
        return hit[1]
    with torch.no_grad():
        ops_ = K.posconv_weight_prep(v.detach(), g.detach().reshape(-1))
    conv.__dict__["_mh_posconv"] = (sig, ops_)
    return ops_


class PosConvFn(torch.autograd.Function):
    """rows [B*T, C] bf16 -> rows + gelu(conv(rows) + bias); implicit-GEMM tcgen05 kernels fwd / dgrad / wgrad."""

    @staticmethod
    def forward(ctx, x, conv, B, T, hook):
        w_fwd, w_bwd, norm = _posconv_operands(conv)
        need = any(ctx.needs_input_grad)
        y, z = K.posconv_fwd(x, w_fwd, conv.bias.detach(), B, T, conv.groups, 128, want_z=need)
        if need:
            ctx.conv, ctx.dims = conv, (B, T)
            ctx.save_for_backward(x, z, w_bwd, norm)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, z, w_bwd, norm = ctx.saved_tensors
        conv = ctx.conv
        B, T = ctx.dims
        dy = dy.contiguous()
        dz = K.gelu_bwd_mul(dy, z)
        dx = K.posconv_dgrad(dz, w_bwd, dy, B, T, conv.groups, 128) if ctx.needs_input_grad[0] else None
        v, g = conv.weight_v, conv.weight_g
        if v.requires_grad or g.requires_grad:
            dw = torch.zeros(v.shape, device=v.device, dtype=torch.float32)
            K.posconv_wgrad(dz, x, dw, B, T, conv.groups, 128)
            K.posconv_weight_bwd(dw, v.detach(), g.detach().reshape(-1), norm, _grad_of(v), _grad_of(g).view(-1))
        if conv.bias is not None and conv.bias.requires_grad:
            K.colsum_add(dz, _grad_of(conv.bias))
        return (dx,) + (None,) * (len(ctx.needs_input_grad) - 1)


def pos_conv(x, conv, B, T):
    return PosConvFn.apply(x, conv, B, T, grad_hook(conv, x.device))


# ---------------------------------------------------------------------------------------------
# fused transformer encoder layer (module.py:82-133 + multihead_attention.py:98-172 +
# forward_multihead_attention.py:39-243)
# ---------------------------------------------------------------------------------------------
SITE_ATTN, SITE_DROP1, SITE_DROP2, SITE_DROP3 = 0, 1, 2, 3


class EncoderLayerFn(torch.autograd.Function):
    """x [B*T, C] bf16 -> layer output [B*T, C] bf16 (post-LN or pre-LN block)."""

    @staticmethod
    def forward(ctx, x, kv_len, layer, B, T, seed, site_base, causal, hook):
        mha = layer.self_attn
        heads = mha.num_heads
        training = layer.training
        p_res = layer.dropout1.p if training else 0.0
        p_act = layer.dropout2.p if training else 0.0
        p_att = mha.dropout_module.p if (training or mha.dropout_module.apply_during_inference) else 0.0
        pre_ln = layer.layer_norm_first
        ln1, ln2 = layer.self_attn_layer_norm, layer.final_layer_norm
        M, C = x.shape
        dev = x.device
        wqkv, bqkv = packed_operands(mha, "qkv", [mha.q_proj, mha.k_proj, mha.v_proj])
        wo, bo = packed_operands(mha, "out", [mha.out_proj])
        w1, b1 = packed_operands(layer, "fc1", [layer.fc1])
        w2, b2 = packed_operands(layer, "fc2", [layer.fc2])
        E, F = wo.shape[1], w1.shape[0]

        mean0 = rstd0 = None
        a_in = x
        if pre_ln:
            a_in, mean0, rstd0 = K.layernorm_fwd(x, ln1.weight.detach(), ln1.bias.detach(), ln1.eps)
        qkv = torch.empty(M, 3 * E, device=dev, dtype=bf16)
        K.gemm(a_in, wqkv, qkv, bias=bqkv)
        ctxv, lse, keep = K.attn_fwd(qkv, kv_len, B, T, heads, causal=causal, p_drop=p_att, seed=seed, site=site_base + SITE_ATTN)
        y1 = torch.empty(M, C, device=dev, dtype=bf16)
        K.gemm(ctxv, wo, y1, epilogue=K.EPI_RES, bias=bo, aux_in=x, p_drop=p_res, seed=seed, site=site_base + SITE_DROP1)
        if pre_ln:
            f_in, mean1, rstd1 = K.layernorm_fwd(y1, ln2.weight.detach(), ln2.bias.detach(), ln2.eps)
            res2 = y1
        else:
            f_in, mean1, rstd1 = K.layernorm_fwd(y1, ln1.weight.detach(), ln1.bias.detach(), ln1.eps)
            res2 = f_in
        pre = torch.empty(M, F, device=dev, dtype=bf16)
        u = torch.empty(M, F, device=dev, dtype=bf16)
        K.gemm(f_in, w1, u, epilogue=K.EPI_GELU, bias=b1, aux_out=pre, p_drop=p_act, seed=seed, site=site_base + SITE_DROP2)
        y2 = torch.empty(M, C, device=dev, dtype=bf16)
        K.gemm(u, w2, y2, epilogue=K.EPI_RES, bias=b2, aux_in=res2, p_drop=p_res, seed=seed, site=site_base + SITE_DROP3)
        if pre_ln:
            out, mean2, rstd2 = y2, None, None
        else:
            out, mean2, rstd2 = K.layernorm_fwd(y2, ln2.weight.detach(), ln2.bias.detach(), ln2.eps)

        if any(ctx.needs_input_grad):
            ctx.layer = layer
            ctx.meta = (B, T, heads, seed, site_base, causal, p_res, p_act, p_att, pre_ln)
            ctx.save_for_backward(x, kv_len, a_in, qkv, ctxv, lse, y1, mean0, rstd0, mean1, rstd1, f_in, pre, u, y2, mean2,
                                  rstd2, wqkv, wo, w1, w2, keep)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x, kv_len, a_in, qkv, ctxv, lse, y1, mean0, rstd0, mean1, rstd1, f_in, pre, u, y2, mean2, rstd2, wqkv, wo, w1,
         w2, keep) = ctx.saved_tensors
        layer = ctx.layer
        mha = layer.self_attn
        B, T, heads, seed, site_base, causal, p_res, p_act, p_att, pre_ln = ctx.meta
        ln1, ln2 = layer.self_attn_layer_norm, layer.final_layer_norm
        E = wo.shape[1]
        dout = dout.contiguous()
        s1, s2, s3 = site_base + SITE_DROP1, site_base + SITE_DROP2, site_base + SITE_DROP3

        # the attention backward's fp32 dQ workspace is zeroed on a side stream NOW, under the FFN backward's GEMMs
        # (a 74 MB memset in front of the attention kernel otherwise: 14 us per layer on the critical path)
        dq_acc, zero_done = _dq_workspace(qkv.device, B * T * E)
        # ---- FFN block ----
        if pre_ln:
            # out = y1 + drop3(fc2(u));  dz2 = dout * mask3, residual gradient = dout
            dy2 = dout
            dz2 = K.dropout_apply(dout, p_res, seed, s3) if p_res > 0.0 else dout
        else:
            b2g = _bias_grad_target(layer.fc2)  # fc2's bias gradient = column sums of dz2: accumulated by the LN backward
            dy2, dz2 = K.layernorm_bwd(dout, y2, ln2.weight.detach(), mean2, rstd2, _grad_of(ln2.weight), _grad_of(ln2.bias),
                                       want_drop=p_res > 0.0, p_out=p_res, seed_out=seed, site_out=s3, colsum_out=b2g)
            if dz2 is None:
                dz2 = dy2
        _wgrad(dz2, u, layer.fc2, bias_done=not pre_ln and b2g is not None)
        dpre = torch.empty_like(pre)
        K.gemm(dz2, w2, dpre, b_mn=True, epilogue=K.EPI_DGELU, aux_in=pre, p_drop=p_act, seed=seed, site=s2)
        _wgrad(dpre, f_in, layer.fc1)
        if pre_ln:
            # f_in = LN2(y1): d y1 = dy2 (residual) + LN2_bwd(dpre W1)
            df = torch.empty_like(f_in)
            K.gemm(dpre, w1, df, b_mn=True)
            dln, _ = K.layernorm_bwd(df, y1, ln2.weight.detach(), mean1, rstd1, _grad_of(ln2.weight), _grad_of(ln2.bias))
            dy1 = dln + dy2
            dz1 = K.dropout_apply(dy1, p_res, seed, s1) if p_res > 0.0 else dy1
        else:
            dx1 = torch.empty_like(f_in)
            K.gemm(dpre, w1, dx1, b_mn=True, epilogue=K.EPI_ADD, aux_in=dy2)
            bog = _bias_grad_target(mha.out_proj)
            dy1, dz1 = K.layernorm_bwd(dx1, y1, ln1.weight.detach(), mean1, rstd1, _grad_of(ln1.weight), _grad_of(ln1.bias),
                                       want_drop=p_res > 0.0, p_out=p_res, seed_out=seed, site_out=s1, colsum_out=bog)
            if dz1 is None:
                dz1 = dy1
        # ---- attention block ----
        _wgrad(dz1, ctxv, mha.out_proj, bias_done=not pre_ln and bog is not None)
        dctx = torch.empty_like(ctxv)
        # out_proj dgrad; its epilogue also emits delta = rowsum(dO * O) per head (64-column epilogue groups = heads)
        delta = torch.empty(B, heads, T, device=dctx.device, dtype=torch.float32)
        K.gemm(dz1, wo, dctx, b_mn=True, epilogue=K.EPI_DELTA, aux_in=ctxv, delta=delta, delta_T=T)
        torch.cuda.current_stream().wait_event(zero_done)
        bqkv = _qkv_bias_grad_target(mha, E)
        dqkv = K.attn_bwd(qkv, kv_len, ctxv, dctx, lse, keep, B, T, heads, causal=causal, p_drop=p_att, seed=seed,
                          site=site_base + SITE_ATTN, dq_acc=dq_acc, delta=delta, bias_grad=bqkv)
        _wgrad_qkv(dqkv, a_in, mha, E, bias_done=bqkv is not None)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if pre_ln:
                da = torch.empty_like(x)
                K.gemm(dqkv, wqkv, da, b_mn=True)
                dln, _ = K.layernorm_bwd(da, x, ln1.weight.detach(), mean0, rstd0, _grad_of(ln1.weight), _grad_of(ln1.bias))
                dx = dln + dy1
            else:
                K.gemm(dqkv, wqkv, dx, b_mn=True, epilogue=K.EPI_ADD, aux_in=dy1)
        hook = getattr(layer, "_mh_grad_ready_hook", None)
        if hook is not None:
            hook(layer)
        return (dx,) + (None,) * (len(ctx.needs_input_grad) - 1)


_DQ_WS = {}


def _dq_workspace(device, numel):
    """Persistent fp32 dQ workspace of the attention backward (one per device and size: the layers of a backward pass
    use it one after the other), zeroed on a side stream; returns (workspace, event that marks the zeroing)."""
    key = (device, numel)
    ws = _DQ_WS.get(key)
    if ws is None:
        ws = _DQ_WS[key] = (torch.empty(numel, device=device, dtype=torch.float32), torch.cuda.Stream(device=device))
    buf, side = ws
    side.wait_stream(torch.cuda.current_stream())  # the previous user (the layer above, or the last step) is done with it
    with torch.cuda.stream(side):
        buf.zero_()
        ev = torch.cuda.Event()
        ev.record(side)
    return buf, ev


def encoder_layer(x, kv_len, layer, B, T, seed, site_base, causal=False):
    return EncoderLayerFn.apply(x, kv_len, layer, B, T, seed, site_base, causal, grad_hook(layer, x.device))


# ---------------------------------------------------------------------------------------------
# dtype / layout boundary
# ---------------------------------------------------------------------------------------------
class MaskRowsToBf16(torch.autograd.Function):
    """bf16(x) with selected rows zeroed (model.py:80 masking, module.py:226-227 padding)."""

    @staticmethod
    def forward(ctx, x, zero_row):
        ctx.save_for_backward(zero_row)
        return K.mask_rows_to_bf16(x, zero_row)

    @staticmethod
    def backward(ctx, dy):
        (zero_row,) = ctx.saved_tensors
        g = K.to_f32(dy)
        if zero_row is not None:
            g = g.masked_fill(zero_row.bool().unsqueeze(1), 0.0)
        return g, None


class ZeroRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, zero_row):
        ctx.save_for_backward(zero_row)
        ctx.mark_dirty(x)
        K.zero_rows_(x, zero_row)
        return x

    @staticmethod
    def backward(ctx, dy):
        (zero_row,) = ctx.saved_tensors
        dy = dy.contiguous().clone()
        K.zero_rows_(dy, zero_row)
        return dy, None


class ToF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return K.to_f32(x)

    @staticmethod
    def backward(ctx, dy):
        return K.to_bf16(dy)


class GatherRows(torch.autograd.Function):
    """hidden[masked_indices] (model.py:148): rows listed in ``idx`` (row-major (b, t) order)."""

    @staticmethod
    def forward(ctx, x, idx, n_idx):
        ctx.save_for_backward(idx)
        ctx.shape, ctx.n_idx = x.shape, n_idx
        return K.gather_rows(x, idx, n_idx)

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        dx = torch.zeros(ctx.shape, device=dy.device, dtype=bf16)
        K.scatter_rows_add_(dy.contiguous(), idx, ctx.n_idx, dx)
        return dx, None, None


# ---------------------------------------------------------------------------------------------
# criteria
# ---------------------------------------------------------------------------------------------
class CrossEntropyFn(torch.autograd.Function):
    """CrossEntropyLoss(ignore_index=-100, reduction='mean') on bf16 logits, fused fwd + bwd.

    ``n_valid`` (device int32 [1], optional) limits the rows that count -- lets a statically
    shaped, CUDA-graph-captured step use a padded row list.  ``reduce_fn`` (optional) is
    applied to the device accumulator [sum_loss, count] before the mean is formed: a
    data-parallel wrapper passes an all-reduce there so the normaliser is the *global*
    masked-frame count, matching nn.DataParallel's gather-then-mean semantics."""

    @staticmethod
    def forward(ctx, logits, labels, weight, n_valid, reduce_fn):
        acc = torch.zeros(2, device=logits.device, dtype=torch.float32)
        K.ce_fwd(logits, labels, acc, n_valid=n_valid)
        local = acc.clone()
        if reduce_fn is not None:
            reduce_fn(acc)
        out = torch.empty(2, device=logits.device, dtype=torch.float32)  # [loss, grad_scale]
        K.ce_finalize(acc, weight, out[0:1], out[1:2])
        ctx.save_for_backward(logits, labels, n_valid, out)
        ctx.local_acc = local
        return out[0]

    @staticmethod
    def backward(ctx, dloss):
        logits, labels, n_valid, out = ctx.saved_tensors
        gs = out[1:2] * dloss.reshape(1).to(torch.float32)
        return K.ce_bwd(logits, labels, gs, n_valid=n_valid), None, None, None, None


def cross_entropy(logits, labels, weight=1.0, n_valid=None, reduce_fn=None):
    return CrossEntropyFn.apply(logits, labels, float(weight), n_valid, reduce_fn)


class KDLossFn(torch.autograd.Function):
    """(1-alpha) * CE(student, label) + alpha * KLDiv_batchmean(log_softmax(s/T), softmax(t/T))
    (distillation/pretrain_expert.py:83-92).  Returns [total, hard, soft, teacher_ce]."""

    @staticmethod
    def forward(ctx, s_logits, t_logits, labels, T, alpha, n_valid, reduce_fn):
        acc = torch.zeros(5, device=s_logits.device, dtype=torch.float32)
        K.kd_fwd(s_logits, t_logits, labels, T, acc, n_valid=n_valid)
        if reduce_fn is not None:
            reduce_fn(acc)
        out = torch.empty(6, device=s_logits.device, dtype=torch.float32)
        K.kd_finalize(acc, alpha, out[0:4], out[4:5], out[5:6])
        ctx.save_for_backward(s_logits, t_logits, labels, n_valid, out)
        ctx.T = T
        terms = out[0:4].clone()  # detached diagnostics: must not keep the autograd graph (and its
        ctx.mark_non_differentiable(terms)  # stream-bound grad accumulators) alive across steps
        return out[0].clone(), terms

    @staticmethod
    def backward(ctx, dloss, _):
        s_logits, t_logits, labels, n_valid, out = ctx.saved_tensors
        d = dloss.reshape(1).to(torch.float32)
        return (K.kd_bwd(s_logits, t_logits, labels, ctx.T, out[4:5] * d, out[5:6] * d, n_valid=n_valid), None, None, None,
                None, None, None)


def kd_loss(s_logits, t_logits, labels, T=1.0, alpha=0.5, n_valid=None, reduce_fn=None):
    return KDLossFn.apply(s_logits, t_logits, labels, float(T), float(alpha), n_valid, reduce_fn)


class L1CosFn(torch.autograd.Function):
    """mean |p - t| + w_cos * mean(-logsigmoid(cos(p, t))) over rows of bf16 [rows, C]."""

    @staticmethod
    def forward(ctx, pred, target, cos_weight):
        rows, cols = pred.shape
        acc = torch.zeros(2, device=pred.device, dtype=torch.float32)
        K.l1cos_fwd(pred, target, acc)
        ctx.save_for_backward(pred, target)
        ctx.w = (1.0 / (rows * cols), cos_weight / rows)
        return acc[0] * ctx.w[0] + acc[1] * ctx.w[1]

    @staticmethod
    def backward(ctx, dloss):
        pred, target = ctx.saved_tensors
        d = dloss.reshape(1).to(torch.float32)
        return K.l1cos_bwd(pred, target, d * ctx.w[0], d * ctx.w[1]), None, None


def l1_cosine_loss(pred, target, cos_weight=1.0):
    return L1CosFn.apply(pred, target, float(cos_weight))
