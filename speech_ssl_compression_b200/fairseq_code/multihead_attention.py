"""Self-attention module with the reference's parameter layout (separate q/k/v/out
``nn.Linear`` leaves that the pruning tools slice and re-parametrise) whose math runs on the
fused kernels: one tcgen05 GEMM for the stacked QKV projection, the flash-style attention
kernel, one GEMM for the output projection.

Mirrors reference ``fairseq_code/multihead_attention.py:24-172``; the functional part
(``pytorch_code/forward_multihead_attention.py:113-243``) has no Python counterpart here.
Inside the encoder the whole layer is one fused autograd function (``ops.EncoderLayerFn``);
``forward`` below is the stand-alone module entry point kept for API compatibility.
"""
import math

import torch
import torch.nn as nn

from .. import kernels as K
from .. import ops


class FairseqDropout(nn.Module):
    """Holds the attention-probability dropout rate (reference fairseq_dropout.py:22-32); the
    dropout itself happens inside the attention kernel."""

    def __init__(self, p):
        super().__init__()
        self.p = p
        self.apply_during_inference = False

    def forward(self, x, inplace: bool = False):
        if self.p > 0 and (self.training or self.apply_during_inference):
            return nn.functional.dropout(x, p=self.p, training=True, inplace=inplace)
        return x


class _SelfAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kv_len, mha, B, T, seed, site, causal, hook):
        heads = mha.num_heads
        p_att = mha.dropout_module.p if (mha.training or mha.dropout_module.apply_during_inference) else 0.0
        wqkv, bqkv = ops.packed_operands(mha, "qkv", [mha.q_proj, mha.k_proj, mha.v_proj])
        wo, bo = ops.packed_operands(mha, "out", [mha.out_proj])
        E = wo.shape[1]
        qkv = torch.empty(x.shape[0], 3 * E, device=x.device, dtype=torch.bfloat16)
        K.gemm(x, wqkv, qkv, bias=bqkv)
        ctxv, lse, keep = K.attn_fwd(qkv, kv_len, B, T, heads, causal=causal, p_drop=p_att, seed=seed, site=site)
        out = torch.empty(x.shape[0], wo.shape[0], device=x.device, dtype=torch.bfloat16)
        K.gemm(ctxv, wo, out, bias=bo)
        if any(ctx.needs_input_grad):
            ctx.mha, ctx.meta = mha, (B, T, heads, seed, site, causal, p_att)
            ctx.save_for_backward(x, kv_len, qkv, ctxv, lse, wqkv, wo, keep)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, kv_len, qkv, ctxv, lse, wqkv, wo, keep = ctx.saved_tensors
        mha = ctx.mha
        B, T, heads, seed, site, causal, p_att = ctx.meta
        E = wo.shape[1]
        dout = dout.contiguous()
        ops._wgrad(dout, ctxv, mha.out_proj)
        dctx = torch.empty_like(ctxv)
        K.gemm(dout, wo, dctx, b_mn=True)
        dqkv = K.attn_bwd(qkv, kv_len, ctxv, dctx, lse, keep, B, T, heads, causal=causal, p_drop=p_att, seed=seed, site=site)
        for i, lin in enumerate((mha.q_proj, mha.k_proj, mha.v_proj)):
            ops._wgrad(dqkv, x, lin, col0=i * E, ncols=E)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            K.gemm(dqkv, wqkv, dx, b_mn=True)
        return (dx,) + (None,) * (len(ctx.needs_input_grad) - 1)


class MultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, kdim=None, vdim=None, dropout=0.0, bias=True, self_attention=False,
                 encoder_decoder_attention=False):
        super().__init__()
        self.embed_dim = embed_dim
        self.kdim = embed_dim if kdim is None else kdim
        self.vdim = embed_dim if vdim is None else vdim
        self.qkv_same_dim = self.kdim == embed_dim and self.vdim == embed_dim
        self.num_heads = num_heads
        self.dropout_module = FairseqDropout(dropout)
        self.head_dim = embed_dim // num_heads
        if self.head_dim * num_heads != embed_dim:
            raise AssertionError("embed_dim must be divisible by num_heads")
        if self.head_dim != 64:
            raise NotImplementedError("the sm_100a attention kernel is specialised for head_dim = 64")
        self.scaling = self.head_dim ** -0.5
        self.self_attention = self_attention
        self.encoder_decoder_attention = encoder_decoder_attention
        if not (self_attention and self.qkv_same_dim):
            raise NotImplementedError("only self-attention with equal q/k/v widths is on the MelHuBERT path")
        # creation order k, v, q, out matters: it fixes the RNG stream of the random init
        self.k_proj = nn.Linear(self.kdim, embed_dim, bias=bias)
        self.v_proj = nn.Linear(self.vdim, embed_dim, bias=bias)
        self.q_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.skip_embed_dim_check = False
        self.need_intermediate = False
        self.context_layer_val = None
        self._rng_calls = 0
        self.reset_parameters()

    def _set_skip_embed_dim_check(self):
        self.skip_embed_dim_check = True

    def _set_need_intermediate(self, state: bool = False):
        self.need_intermediate = state

    def reset_parameters(self):
        gain = 1 / math.sqrt(2) if self.qkv_same_dim else 1.0
        for proj in (self.k_proj, self.v_proj, self.q_proj):
            nn.init.xavier_uniform_(proj.weight, gain=gain)
        nn.init.xavier_uniform_(self.out_proj.weight)
        if self.out_proj.bias is not None:
            nn.init.constant_(self.out_proj.bias, 0.0)

    def forward(self, query, key=None, value=None, key_padding_mask=None, need_weights=False, attn_mask=None,
                valid_lens=None):
        """query: (T, B, C) like the reference.  ``key_padding_mask`` (B, T) bool with True at
        padded keys -- suffix padding (what the datasets produce); or pass ``valid_lens``."""
        if need_weights or self.need_intermediate:
            raise NotImplementedError("attention weights / per-head context are never materialised by the fused kernel")
        if (key is not None and key is not query) or (value is not None and value is not query):
            raise NotImplementedError("only self-attention is supported")
        T, B, C = query.shape
        if valid_lens is None:
            valid_lens = (T - key_padding_mask.sum(dim=1)).to(torch.int32) if key_padding_mask is not None else None
        x = query.transpose(0, 1).reshape(B * T, C)
        x = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
        self._rng_calls += 1
        out = _SelfAttentionFn.apply(x.contiguous(), valid_lens, self, B, T, self._rng_calls, 0, attn_mask is not None,
                                     ops.grad_hook(self, x.device))
        return out.view(B, T, -1).transpose(0, 1).to(query.dtype), None
