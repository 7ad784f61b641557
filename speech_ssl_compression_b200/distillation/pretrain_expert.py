"""MelHuBERT distillation expert -- drop-in for reference
``distillation/pretrain_expert.py:11-141`` (identical to
``upstream/melhubert_distiller/pretrain_expert.py`` up to the student's config key).

Teacher forward without autograd (kept in train mode like the reference), student
forward with the teacher's span mask, then the fused KD criterion
``(1-alpha) CE + alpha KL_batchmean`` on the cluster logits.  ``forward`` returns
``(loss, 1)`` -- the reference returns a bare tensor which its own runner cannot unpack (Q2).

Extension named by north_star (not in the reference): ``loss_param.type: l1cos``
distils hidden states layer by layer with the fused L1 + cosine kernel
(``loss_param.cos_weight``, ``loss_param.layer_map`` = {student_layer: teacher_layer}).
"""
import torch
import torch.nn as nn

from .. import kernels as K
from .. import ops
from ..model import MelHuBERTConfig, MelHuBERTModel


class MelHuBERTDistiller(nn.Module):
    STUDENT_KEY = "melhubert"

    def __init__(self, upstream_config, initial_weight=None, device="cuda", multi_gpu=False):
        super().__init__()
        self.initial_weight, self.device, self.multi_gpu = initial_weight, device, multi_gpu
        self.upstream_config = upstream_config
        self._init_model()
        lp = upstream_config["loss_param"]
        self.loss_temp, self.loss_alpha, self.loss_type = lp["T"], lp["alpha"], lp["type"]
        if self.loss_type not in ("masked", "nomasked", "l1cos"):
            raise SystemExit(f"[Distiller] - No such loss type {self.loss_type}")
        self.mask_or_not = self.loss_type == "masked" or (self.loss_type == "l1cos" and lp.get("mask", False))
        self.cos_weight = float(lp.get("cos_weight", 1.0))
        ls, lt = self.student_config.encoder_layers, self.teacher_config.encoder_layers
        self.layer_map = lp.get("layer_map") or {i: (i + 1) * lt // ls - 1 for i in range(ls)}
        self.dp = None
        if multi_gpu:
            from ..parallel import DataParallelB200

            self.dp = DataParallelB200(self.model)
            print("[Distiller] - Multi-GPU training Enabled: " + str(self.dp.world_size))
        print("[Distiller] - Number of parameters: " +
              str(sum(p.numel() for p in self.model.parameters() if p.requires_grad)))
        self.last_terms = None

    def _init_model(self):
        print("[Distiller] - Initializing model...")
        self.student_config = MelHuBERTConfig(self.upstream_config[self.STUDENT_KEY])
        self.model = MelHuBERTModel(self.student_config)
        self.teacher_config = MelHuBERTConfig(self.upstream_config["teacher"])
        self.teacher_model = MelHuBERTModel(self.teacher_config)
        assert self.initial_weight, "Please specify teacher's weight by -i argument"
        states = torch.load(self.initial_weight, map_location="cpu", weights_only=False)
        try:
            self.teacher_model.load_state_dict(states["model"])
            print(f"[Distiller] - Load teacher model's weight from {self.initial_weight}")
        except Exception as e:
            raise NotImplementedError(f"Could not load the teacher model's weight: {e}")
        for p in self.teacher_model.parameters():
            p.requires_grad_(False)  # the reference leaves them trainable-but-gradless (Q7); same optimizer effect
        if self.upstream_config[self.STUDENT_KEY].get("initial_from_teacher", False):
            print("[Distiller] - Initializing from teacher")
            self.model.encoder.pos_conv.load_state_dict(self.teacher_model.encoder.pos_conv.state_dict())
            for l in range(self.student_config.encoder_layers):
                self.model.encoder.layers[l].load_state_dict(self.teacher_model.encoder.layers[l].state_dict())

    def load_model(self, init_ckpt):
        assert "model" in init_ckpt
        self.model.load_state_dict(init_ckpt["model"])
        ops.bump_weight_epoch()

    def add_state_to_save(self, all_states):
        all_states["model"] = self.model.state_dict()
        all_states["Upstream_Config"] = self.upstream_config
        return all_states

    def loss_fn_kd(self, outputs, labels, teacher_outputs, T=1, alpha=0.5):
        """Returns (total, hard, soft, teacher_ce) like reference :83-92 (fused kernel)."""
        reduce_fn = self.dp.all_reduce_sum if self.dp is not None else None
        total, terms = ops.kd_loss(outputs, teacher_outputs, labels, T, alpha, reduce_fn=reduce_fn)
        return total, terms[1], terms[2], terms[3]

    def acc(self, outputs, labels):
        return torch.sum(torch.argmax(outputs, dim=1) == labels).item(), len(labels)

    def forward(self, data, records=None, global_step=0, log_step=1000, **kwargs):
        audio_feat, label, pad_mask, audio_len = data[0], data[1], data[2], data[3]
        audio_feat = audio_feat.to(self.device, non_blocking=True)
        label = label.to(self.device, non_blocking=True)
        pad_mask = pad_mask.to(self.device, non_blocking=True)
        lens = list(audio_len) if audio_len is not None else None
        want_hidden = self.loss_type == "l1cos"
        with torch.no_grad():
            t_out = self.teacher_model(audio_feat, pad_mask, label, mask=self.mask_or_not, get_hidden=want_hidden,
                                       valid_lens=lens)
        s_out = self.model(audio_feat, pad_mask, label, mask=self.mask_or_not, get_hidden=want_hidden,
                           teacher_mask_indices=t_out[7], valid_lens=lens)
        if self.loss_type == "l1cos":
            loss = 0.0
            for s_idx, t_idx in self.layer_map.items():
                s_h, t_h = s_out[5][int(s_idx)], t_out[5][int(t_idx)]
                C = s_h.shape[-1]
                loss = loss + ops.l1_cosine_loss(_rows_bf16(s_h),
                                                 _rows_bf16(t_h.detach()), self.cos_weight)
            loss = loss / max(len(self.layer_map), 1)
            if self.dp is not None and self.dp.enabled:
                # a per-rank frame mean: the bucket all-reduce SUMS gradients over ranks, so the 1 / world that makes
                # them the global mean (CE / KD get it through their global frame count) is applied here
                loss = loss / self.dp.world_size
            return loss, 1
        if self.loss_type == "masked":
            total, h, s, t = self.loss_fn_kd(s_out[1], s_out[3], t_out[1], T=self.loss_temp, alpha=self.loss_alpha)
        else:
            total, h, s, t = self.loss_fn_kd(s_out[2], s_out[4], t_out[2], T=self.loss_temp, alpha=self.loss_alpha)
        self.last_terms = (h.detach(), s.detach(), t.detach())
        return total, 1


class _RowsBf16(torch.autograd.Function):
    """(B, T, C) fp32 hidden -> bf16 [B*T, C] rows for the criterion kernels."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = x.shape
        return K.to_bf16(x.reshape(-1, x.shape[-1]))

    @staticmethod
    def backward(ctx, dy):
        return K.to_f32(dy).view(ctx.shape)


def _rows_bf16(x):
    return _RowsBf16.apply(x)
