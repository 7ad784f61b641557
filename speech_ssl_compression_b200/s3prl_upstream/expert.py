"""The "upstream expert API" of the reference's S3PRL packaging (``s3prl_upstream/expert.py:46-139``) on the
B200 path: waveforms in, ``{"hidden_states": [pre_feat] + layer_hiddens, "last_hidden_state": hidden}`` out.

Same constructor arguments, checkpoint handling (``Pruned_heads`` -> shrunk attention layers, ``Pruning`` ->
Identity masks, load, ``prune.remove``) and ``get_downsample_rates`` as the reference class; the differences are the
ones the missing ``s3prl`` package forces (plain ``nn.Module`` instead of ``UpstreamBase``) and the point of this
repo: the log-mel front-end runs in ``mh_fbank`` on the GPU the waveforms live on (``frontend/fbank.py``) instead
of torchaudio, and the encoder is the sm_100a kernel path.  CPU waveforms are refused (no CPU fallback).
"""
import numpy as np
import torch
import torch.nn as nn

from ..frontend.fbank import kaldi_fbank, stack_frames
from ..model import MelHuBERTConfig, MelHuBERTModel
from ..surgery import apply_pruned_heads_record


def load_mean_std(mean_std_npy_path):
    """(2, 40) float64 file: row 0 mean, row 1 std (s3prl_upstream/expert.py:17-21)."""
    ms = np.load(mean_std_npy_path)
    return torch.Tensor(ms[0].reshape(-1)), torch.Tensor(ms[1].reshape(-1))


class UpstreamExpert(nn.Module):
    def __init__(self, ckpt, mode, fp, mean_std_npy_path, model_config=None, **kwargs):
        super().__init__()
        self.mode, self.fp = mode, fp
        states = torch.load(ckpt, map_location="cpu", weights_only=False)
        up = states["Upstream_Config"]
        cfg = MelHuBERTConfig(up["melhubert"] if "melhubert" in up else up["hubert"])
        self.upstream_model = MelHuBERTModel(cfg)
        if "Pruned_heads" in states:   # head-pruned: layers shrink to their kept-head count before loading
            apply_pruned_heads_record(self.upstream_model, states["Pruned_heads"])
        params = None
        if "Pruning" in states:        # weight-pruned: *_orig / *_mask keys need the reparametrisation in place
            from ..pytorch_code import prune
            from ..weight_pruning.wp_utils import get_params_to_prune

            params, _ = get_params_to_prune(self.upstream_model)
            prune.global_unstructured(params, pruning_method=prune.Identity)
        self.upstream_model.load_state_dict(states["model"])
        if params is not None:
            for module, name in params:
                prune.remove(module, name)
        self.mean, self.std = load_mean_std(mean_std_npy_path)

    def get_downsample_rates(self, key: str) -> int:
        return {20: 320, 10: 160}[self.fp]

    def forward(self, wavs, no_pred=True, norm=True):
        """wavs: list of 1-D float waveforms (16 kHz, [-1, 1)) on the model's CUDA device."""
        dev = wavs[0].device
        if dev.type != "cuda":
            raise RuntimeError("UpstreamExpert: waveforms must live on the CUDA device (there is no CPU path)")
        feat, frames = kaldi_fbank(wavs, self.mean.numpy(), self.std.numpy(), device=dev)
        mel, lens = stack_frames(feat, frames, self.fp)
        mel = mel[:, : max(lens)].contiguous()
        pad = (torch.arange(mel.shape[1], device=dev)[None, :] < torch.tensor(lens, device=dev)[:, None]).float()
        out = self.upstream_model(mel, pad, mask=False, no_pred=True, get_hidden=True)
        hidden, layer_hiddens, pre_feat = out[0], out[5], out[6]   # (7-tuple with no_pred, model.py:143)
        return {"hidden_states": [pre_feat] + layer_hiddens, "last_hidden_state": hidden}
