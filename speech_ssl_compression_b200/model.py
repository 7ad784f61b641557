"""MelHuBERT model -- drop-in for reference ``model.py:17-163`` (``MelHuBERTModel``) with the
encoder hot path on hand-written sm_100a kernels.

Same constructor argument, module tree / state_dict names, forward signature and return
tuple.  Differences a caller can observe (all listed in DESIGN.md):
  * compute precision is bf16 with fp32 accumulation (the reference trains under fp16
    autocast); ``hidden`` / ``layer_hiddens`` / ``pre_feat`` are returned as fp32, logits as bf16;
  * CUDA only -- there is no CPU path;
  * key padding must be suffix padding (what both datasets and ``extract_feature.py`` produce).
"""
import numpy as np
import torch
from torch import nn

from . import kernels as K
from . import ops
from .fairseq_code import compute_mask_indices
from .model_config import MelHuBERTConfig
from .module import TransformerEncoder

_INSTANCES = [0]


class MelHuBERTModel(nn.Module):
    def __init__(self, model_config: MelHuBERTConfig):
        super().__init__()
        self.model_config = model_config
        self.n_encoder_layers = model_config.encoder_layers
        print(f"[MelHuBERTModel] - Encoder layer = {self.n_encoder_layers}")
        self.pre_extract_proj = (nn.Linear(model_config.feat_emb_dim, model_config.encoder_embed_dim)
                                 if model_config.feat_emb_dim != model_config.encoder_embed_dim else None)
        if model_config.encoder_layers <= 0:
            raise NotImplementedError("encoder_layers = 0 (GELU-only 'encoder') is outside the MelHuBERT hot path")
        self.encoder = TransformerEncoder(model_config)
        if model_config.learnable_mask_emb:
            dim = model_config.feat_emb_dim if model_config.mask_before_proj else model_config.encoder_embed_dim
            self.mask_emb = nn.Parameter(torch.FloatTensor(dim).uniform_())
        else:
            self.mask_emb = 0
        self.final_proj = nn.Linear(model_config.encoder_embed_dim, model_config.num_cluster)
        _INSTANCES[0] += 1
        self._seed_base = (torch.initial_seed() * 1000003 + _INSTANCES[0] * 7919) & ((1 << 62) - 1)
        self._calls = 0
        # static_rows: logits are produced for every frame position (unselected rows get label
        # -100) so that all shapes are independent of the random mask -> CUDA-graph capturable.
        self.static_rows = False

    # ------------------------------------------------------------------------------------------
    def _draw_mask(self, B, T, lens, padding_mask, teacher_mask_indices, device):
        cfg = self.model_config
        if cfg.mask_prob <= 0:
            return None
        if teacher_mask_indices is not None:
            return teacher_mask_indices.to(device)
        m = compute_mask_indices((B, T), padding_mask if lens is None else None, cfg.mask_prob, cfg.mask_length,
                                 cfg.mask_selection, cfg.mask_other, min_masks=2, no_overlap=cfg.no_mask_overlap,
                                 min_space=cfg.mask_min_space, require_same_masks=False, valid_lens=lens)
        return torch.from_numpy(m).to(device, non_blocking=True)

    def _apply_mask(self, x3d, mask_indices):
        """x[mask] = mask_emb, in place on the caller's tensor like the reference (model.py:80)."""
        if mask_indices is None:
            return x3d
        if isinstance(self.mask_emb, nn.Parameter):
            x3d = x3d.clone()
            x3d[mask_indices] = self.mask_emb.to(x3d.dtype)
            return x3d
        x3d.masked_fill_(mask_indices.unsqueeze(-1), float(self.mask_emb))
        return x3d

    def forward(self, feat, pad_mask, cluster_label=None, no_pred=False, mask=False, get_hidden=False,
                teacher_mask_indices=None, valid_lens=None):
        """feat (B, T, D) fp32 on the GPU; pad_mask (B, T) float/bool, 1 = valid frame.
        ``valid_lens`` (optional host list) avoids a device sync when the caller knows the lengths.
        Returns the reference 8-tuple (7-tuple with ``no_pred``)."""
        cfg = self.model_config
        if not feat.is_cuda:
            raise RuntimeError("MelHuBERTModel (B200 build) runs on CUDA tensors only; there is no CPU fallback")
        B, T, D = feat.shape
        dev = feat.device
        valid = pad_mask.bool()
        padded = ~valid
        kv_len = valid.sum(dim=1).to(torch.int32)
        pad_rows = padded.reshape(-1).to(torch.uint8)
        self._calls += 1
        seed = (self._seed_base + self._calls * 104729) & ((1 << 62) - 1)

        mask_indices = None
        if mask and cfg.mask_before_proj:
            mask_indices = self._draw_mask(B, T, valid_lens, padded, teacher_mask_indices, dev)
            feat = self._apply_mask(feat, mask_indices)
        x = ops.MaskRowsToBf16.apply(feat.reshape(B * T, D), None)
        pre = ops.linear(x, self.pre_extract_proj) if self.pre_extract_proj is not None else x
        if mask and not cfg.mask_before_proj:
            mask_indices = self._draw_mask(B, T, valid_lens, padded, teacher_mask_indices, dev)
            if mask_indices is not None:
                if isinstance(self.mask_emb, nn.Parameter):
                    pre = self._apply_mask(pre.view(B, T, -1), mask_indices).reshape(B * T, -1)
                else:
                    pre = ops.ZeroRows.apply(pre, mask_indices.reshape(-1).to(torch.uint8))
        if mask_indices is None:
            mask_indices = torch.zeros(B, T, dtype=torch.bool, device=dev)

        causal = cfg.attention_type == "causal"
        hidden, hiddens = self.encoder.forward_rows(pre, pad_rows, kv_len, B, T, seed, causal, get_hidden)
        C = hidden.shape[1]
        to3d = lambda t: ops.ToF32.apply(t).view(B, T, -1)  # noqa: E731
        hidden_out = to3d(hidden)
        layer_hiddens = [to3d(h) for h in hiddens[:-1]] + ([hidden_out] if hiddens else [])
        pre_feat = to3d(pre)
        if no_pred:
            return hidden_out, None, None, None, None, layer_hiddens, pre_feat
        assert cluster_label is not None
        label_rows = cluster_label.reshape(-1)

        def predict(select):
            sel = select.reshape(-1).to(torch.uint8)
            idx, count = K.select_rows(sel)
            n = B * T if self.static_rows else int(count.item())
            if n == 0:  # e.g. the masked set of an un-masked forward: (0, K) logits like the reference
                return (hidden.new_zeros((0, self.final_proj.out_features)),
                        torch.empty(0, dtype=torch.int64, device=dev))
            rows = ops.GatherRows.apply(hidden, idx, n)
            logits = ops.linear(rows, self.final_proj)
            return logits, K.gather_labels(label_rows, idx, n)

        logit_m = label_m = logit_u = label_u = None
        if not cfg.skip_masked:
            logit_m, label_m = predict(valid & mask_indices)
        if not cfg.skip_nomask:
            logit_u, label_u = predict(valid & ~mask_indices)
        return hidden_out, logit_m, logit_u, label_m, label_u, layer_hiddens, pre_feat, mask_indices
