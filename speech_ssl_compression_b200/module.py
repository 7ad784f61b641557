"""Transformer encoder of MelHuBERT on the fused sm_100a kernels.

Module tree, parameter names, construction order (hence random-init RNG stream) and call
signatures follow reference ``module.py:17-257``; the math does not run through those
modules' ``forward`` -- each layer is ONE autograd function (``ops.EncoderLayerFn``) issuing
four tcgen05 GEMMs with fused epilogues, the flash-style attention kernel and two LayerNorm
kernels over bf16 ``[B*T, C]`` activations.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .fairseq_code import MultiheadAttention, init_bert_params

SITE_ENCODER_DROPOUT = 0xFFF0


class SamePad(nn.Module):
    """Drops the trailing frame an even-sized 'same' convolution produces (same_pad.py:17-28)."""

    def __init__(self, kernel_size, causal=False):
        super().__init__()
        self.remove = kernel_size - 1 if causal else (1 if kernel_size % 2 == 0 else 0)

    def forward(self, x):
        return x[:, :, : -self.remove] if self.remove > 0 else x


class TransformerSentenceEncoderLayer(nn.Module):
    def __init__(self, embedding_dim=768, ffn_embedding_dim=3072, num_attention_heads=8, dropout=0.1,
                 attention_dropout=0.1, activation_dropout=0.1, activation_fn="relu", layer_norm_first=False):
        super().__init__()
        if activation_fn != "gelu":
            raise NotImplementedError("the fused FFN epilogue implements erf-GELU (all shipped configs)")
        self.embedding_dim = embedding_dim
        self.dropout = dropout
        self.activation_dropout = activation_dropout
        # registration order fixes both state_dict order and the init RNG stream
        self.self_attn = MultiheadAttention(embedding_dim, num_attention_heads, dropout=attention_dropout,
                                            self_attention=True)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(activation_dropout)
        self.dropout3 = nn.Dropout(dropout)
        self.layer_norm_first = layer_norm_first
        self.self_attn_layer_norm = nn.LayerNorm(embedding_dim)
        self.fc1 = nn.Linear(embedding_dim, ffn_embedding_dim, bias=True)
        self.fc2 = nn.Linear(ffn_embedding_dim, embedding_dim, bias=True)
        self.final_layer_norm = nn.LayerNorm(embedding_dim)
        self._calls = 0

    def forward_rows(self, x, kv_len, B, T, seed, site_base, causal=False):
        """Fast path: x bf16 [B*T, C] (batch-major rows) -> same shape."""
        return ops.encoder_layer(x, kv_len, self, B, T, seed, site_base, causal)

    def forward(self, x, self_attn_mask=None, self_attn_padding_mask=None, need_weights=False, att_args=None):
        """Reference call shape: x (T, B, C); returns (x, None)."""
        if need_weights:
            raise NotImplementedError("attention weights are never materialised by the fused kernel")
        T, B, C = x.shape
        kv_len = None
        if self_attn_padding_mask is not None:
            kv_len = (T - self_attn_padding_mask.sum(dim=1)).to(torch.int32)
        rows = x.transpose(0, 1).reshape(B * T, C).to(torch.bfloat16).contiguous()
        self._calls += 1
        out = self.forward_rows(rows, kv_len, B, T, self._calls, 0, self_attn_mask is not None)
        return out.view(B, T, C).transpose(0, 1).to(x.dtype), None


class TransformerEncoder(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.dropout = args.dropout
        self.embedding_dim = args.encoder_embed_dim
        self.ffn_embedding_dim = args.encoder_ffn_embed_dim
        self.pos_emb_type = args.pos_emb_type
        if self.pos_emb_type != "conv" or args.pos_conv_depth != 1:
            raise NotImplementedError(f"positional embedding {self.pos_emb_type}/depth {args.pos_conv_depth} "
                                      "is not used by any MelHuBERT config")
        conv = nn.Conv1d(self.embedding_dim, self.embedding_dim, kernel_size=args.conv_pos,
                         padding=args.conv_pos // 2, groups=args.conv_pos_groups)
        nn.init.normal_(conv.weight, mean=0, std=math.sqrt(4.0 / (args.conv_pos * self.embedding_dim)))
        nn.init.constant_(conv.bias, 0)
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            conv = nn.utils.weight_norm(conv, name="weight", dim=2)
        self.pos_conv = nn.Sequential(conv, SamePad(args.conv_pos), nn.GELU())
        self.layers = nn.ModuleList([
            TransformerSentenceEncoderLayer(
                embedding_dim=self.embedding_dim, ffn_embedding_dim=args.encoder_ffn_embed_dim,
                num_attention_heads=args.encoder_attention_heads, dropout=self.dropout,
                attention_dropout=args.attention_dropout, activation_dropout=args.activation_dropout,
                activation_fn=args.activation_fn, layer_norm_first=args.layer_norm_first)
            for _ in range(args.encoder_layers)])
        self.layer_norm_first = args.layer_norm_first
        self.layer_norm = nn.LayerNorm(self.embedding_dim)
        self.layerdrop = args.encoder_layerdrop
        self.apply(init_bert_params)

    # -- fast path ------------------------------------------------------------------------------
    def forward_rows(self, x, pad_rows, kv_len, B, T, seed, causal=False, get_hidden=False):
        """x: bf16 [B*T, C] (modified in place: padded rows are zeroed, as the reference does to
        pre_feat, module.py:226-227).  Returns (rows, [per-layer rows])."""
        C = x.shape[1]
        if pad_rows is not None:
            x = ops.ZeroRows.apply(x, pad_rows)
        if not ops.posconv_supported(self.pos_conv[0]):
            raise NotImplementedError("the positional-conv kernels are built for Conv1d(k=128, pad=64, 48 channels per "
                                      "group, weight-normed) -- the shape of every MelHuBERT config")
        x = ops.pos_conv(x, self.pos_conv[0], B, T)  # x + gelu(same_pad(conv_wn(x)))
        p = self.dropout if self.training else 0.0
        if not self.layer_norm_first:
            x = ops.layer_norm(x, self.layer_norm, p, seed, SITE_ENCODER_DROPOUT)
        elif p > 0:
            x = F.dropout(x, p=p, training=True)
        hiddens = []
        for i, layer in enumerate(self.layers):
            skip_draw = np.random.random()  # drawn every layer, even with layerdrop = 0 (module.py:243)
            if not self.training or skip_draw > self.layerdrop:
                x = layer.forward_rows(x, kv_len, B, T, seed, 8 * (i + 1), causal)
                if get_hidden:
                    hiddens.append(x)
        if self.layer_norm_first:
            x = ops.layer_norm(x, self.layer_norm)
        return x, hiddens

    # -- reference call shape -----------------------------------------------------------------
    def forward(self, x, padding_mask=None, attn_mask=None, get_hidden=False):
        """x (B, T, C) float; padding_mask (B, T) bool, True at padded frames (suffix padding)."""
        B, T, C = x.shape
        kv_len = pad_rows = None
        if padding_mask is not None:
            kv_len = (T - padding_mask.sum(dim=1)).to(torch.int32)
            pad_rows = padding_mask.reshape(-1).to(torch.uint8)
        rows = x.reshape(B * T, C).to(torch.bfloat16).contiguous()
        self._calls = getattr(self, "_calls", 0) + 1
        out, hid = self.forward_rows(rows, pad_rows, kv_len, B, T, self._calls, attn_mask is not None, get_hidden)
        return out.view(B, T, C).to(x.dtype), [h.view(B, T, C).to(x.dtype) for h in hid]
