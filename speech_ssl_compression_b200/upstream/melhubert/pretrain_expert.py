"""MelHuBERT pre-training expert -- drop-in for reference
``upstream/melhubert/pretrain_expert.py:12-121`` (boundary A of SURVEY.md §8b).

``expert(data, global_step=, log_step=) -> (loss, 1)`` with
``data = (feat (B,T,D) f32, label (B,T) i64, pad_mask (B,T) f32, lens)``.  The masked-frame
cross-entropy is the fused criterion kernel; with ``multi_gpu`` the model runs one process
per GPU (``parallel.DataParallelB200``) instead of ``nn.DataParallel`` threads: per-layer
gradient all-reduce overlapped with backward and a 2-scalar all-reduce of (sum CE, count) so
the loss is the mean over the *global* masked-frame set, as DataParallel's gather-then-mean is.
"""
import torch
import torch.nn as nn

from ... import ops
from ...model import MelHuBERTConfig, MelHuBERTModel
from ...pytorch_code import prune
from ...surgery import apply_pruned_heads_record
from ...weight_pruning.wp_utils import get_params_to_prune


class MelHuBERTPretrainer(nn.Module):
    def __init__(self, upstream_config, initial_weight=None, device="cuda", multi_gpu=False, **kwargs):
        super().__init__()
        self.initial_weight = initial_weight
        self.device = device
        self.multi_gpu = multi_gpu
        self.upstream_config = upstream_config
        self.pruned_heads = None
        self._init_model()
        self.dp = None
        if self.multi_gpu:
            from ...parallel import DataParallelB200

            self.dp = DataParallelB200(self.model)
            print("[Pretrainer] - Multi-GPU training Enabled: " + str(self.dp.world_size))
        print("[Pretrainer] - Number of parameters: " +
              str(sum(p.numel() for p in self.model.parameters() if p.requires_grad)))

    def _init_model(self):
        print("[Pretrainer] - Initializing model...")
        self.model_config = MelHuBERTConfig(self.upstream_config["melhubert"])
        self.model = MelHuBERTModel(self.model_config)
        if not self.initial_weight:
            return
        all_states = torch.load(self.initial_weight, map_location="cpu", weights_only=False)
        if "Pruned_heads" in all_states:  # head-pruned ckpt: shrink q/k/v/out first
            self.pruned_heads = all_states["Pruned_heads"]
            apply_pruned_heads_record(self.model, self.pruned_heads)
        if "Pruning" in all_states:  # weight-pruned ckpt: install all-ones masks so *_orig / *_mask load
            params, _ = get_params_to_prune(self.model)
            prune.global_unstructured(params, pruning_method=prune.Identity)
        try:
            self.model.load_state_dict(all_states["model"])
            print(f"[Pretrainer] Load initilization model weight from {self.initial_weight}")
        except Exception as e:
            raise NotImplementedError(f"Could not load the initilization weight: {e}")

    def load_model(self, init_ckpt):
        assert "model" in init_ckpt
        self.model.load_state_dict(init_ckpt["model"])
        ops.bump_weight_epoch()

    def add_state_to_save(self, all_states):
        all_states["model"] = self.model.state_dict()
        all_states["Upstream_Config"] = self.upstream_config
        if self.pruned_heads:
            all_states["Pruned_heads"] = self.pruned_heads
        return all_states

    def forward(self, data, global_step=0, log_step=1000):
        audio_feat, label, pad_mask, audio_len = data[0], data[1], data[2], data[3]
        label = label.to(self.device, non_blocking=True)
        audio_feat = audio_feat.to(self.device, non_blocking=True)
        pad_mask = pad_mask.to(self.device, non_blocking=True)
        lens = list(audio_len) if audio_len is not None else None
        _, logit_m, logit_u, label_m, label_u, _, _, _ = self.model(audio_feat, pad_mask, label, mask=True, valid_lens=lens)
        reduce_fn = self.dp.all_reduce_sum if self.dp is not None else None
        loss = 0.0
        cfg = self.model_config
        if logit_m is not None and cfg.pred_masked_weight > 0:
            loss = loss + ops.cross_entropy(logit_m, label_m, cfg.pred_masked_weight, reduce_fn=reduce_fn)
        if logit_u is not None and cfg.pred_nomask_weight > 0:
            loss = loss + ops.cross_entropy(logit_u, label_u, cfg.pred_nomask_weight, reduce_fn=reduce_fn)
        return loss, 1
