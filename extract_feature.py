#!/usr/bin/env python
"""Feature extraction entry point -- same flags, front-end and checkpoint handling as the reference's
``extract_feature.py:14-149``; the model forward runs on the sm_100a kernels.

    python extract_feature.py -m MODE -c CKPT -f 20 -d 960 [--device cuda] [wav/flac ...]

Front-end (host side, like the reference): FLAC/WAV decode -> x 2^15 -> Kaldi fbank(40 mel, hamming, 25/10 ms)
-> (y - mean) / std -> 20 ms: stack even/odd frames to 80-d -> pad_sequence + pad mask.  ``torchaudio.load``
has no decoder backend in this image, so 16-bit FLAC is decoded by ``frontend/flac.py`` (checked against the
STREAMINFO MD5).  Without ``-c`` a random-init model (seed 1337) is used, as in BASELINE.json config 1.
"""
import argparse
import os
import random
import sys

import numpy as np
import torch
from torch.nn.utils.rnn import pad_sequence

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
EXAMPLE = os.path.join(ROOT, "tests", "golden", "example")


def get_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-m", "--mode", default="melhubert",
                    choices=["melhubert", "weight-pruning", "head-pruning", "row-pruning", "distillation"])
    ap.add_argument("-c", "--checkpoint", help="Path to model checkpoint (default: random init, seed 1337)")
    ap.add_argument("-f", "--fp", type=int, default=20, help="frame period")
    ap.add_argument("-d", "--hours", type=int, choices=[360, 960], default=960)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--host-fbank", action="store_true", help="compute the fbank with torchaudio on the host like the reference")
    ap.add_argument("wavs", nargs="*", help="audio files (default: the two example FLACs)")
    return ap.parse_args(argv)


def load_waveform(path):
    """-> float32 (1, n) in [-1, 1) and the sample rate"""
    if path.lower().endswith(".flac"):
        from speech_ssl_compression_b200.frontend.flac import decode_flac

        pcm, sr, md5_ok, _ = decode_flac(path)
        if not md5_ok:
            raise RuntimeError(f"{path}: decoded PCM does not match the STREAMINFO MD5")
        return torch.from_numpy(pcm.astype(np.float32) / 32768.0).unsqueeze(0), sr
    import torchaudio

    return torchaudio.load(path)


def extract_fbank(path, mean, std, fp=20):
    import torchaudio

    wav, sr = load_waveform(path)
    y = torchaudio.compliance.kaldi.fbank(wav * (2 ** 15), num_mel_bins=40, sample_frequency=16000, window_type="hamming",
                                          frame_length=25, frame_shift=10)
    y = (y - mean) / std
    if fp == 20:
        odd, even = y[::2, :], y[1::2, :]
        if odd.shape[0] != even.shape[0]:
            even = torch.cat((even, torch.zeros(1, even.shape[1])), dim=0)
        y = torch.cat((odd, even), dim=1)
    return y


def prepare_data_gpu(paths, fp=20, hours=360, device="cuda"):
    """Same tuple as prepare_data with the fbank / normalisation / frame stacking on the GPU (``mh_fbank``): only the
    FLAC bitstream decode stays on the host."""
    from speech_ssl_compression_b200.frontend.fbank import kaldi_fbank, stack_frames

    ms = np.load(os.path.join(EXAMPLE, f"libri-{hours}-mean-std.npy"))
    waves = []
    for p in paths:
        wav, sr = load_waveform(p)
        if sr != 16000:
            raise RuntimeError(f"{p}: sample rate {sr}, the front-end is built for 16 kHz")
        waves.append(wav.reshape(-1))
    feat, frames = kaldi_fbank(waves, ms[0].reshape(-1), ms[1].reshape(-1), device=device)
    mel, lens = stack_frames(feat, frames, fp)
    mel = mel[:, : max(lens)].contiguous()
    pad = (torch.arange(mel.shape[1], device=mel.device)[None, :] < torch.tensor(lens, device=mel.device)[:, None]).float()
    return mel, lens, pad


def prepare_data(paths, fp=20, hours=360):
    ms = np.load(os.path.join(EXAMPLE, f"libri-{hours}-mean-std.npy"))
    mean, std = torch.Tensor(ms[0].reshape(-1)), torch.Tensor(ms[1].reshape(-1))
    mels = [extract_fbank(p, mean, std, fp) for p in paths]
    lens = [len(m) for m in mels]
    mel = pad_sequence(mels, batch_first=True)
    pad = torch.ones(mel.shape[:-1])
    for i, l in enumerate(lens):
        pad[i, l:] = 0
    return mel, lens, pad


def load_model(args):
    from speech_ssl_compression_b200.model import MelHuBERTConfig, MelHuBERTModel
    from speech_ssl_compression_b200.pytorch_code import prune
    from speech_ssl_compression_b200.surgery import apply_pruned_heads_record
    from speech_ssl_compression_b200.weight_pruning.wp_utils import get_params_to_prune

    if not args.checkpoint:
        random.seed(1337); np.random.seed(1337); torch.manual_seed(1337)
        cfg = dict(feat_emb_dim=80 if args.fp == 20 else 40, encoder_layers=12, mask_prob=0.7,
                   mask_length=5 if args.fp == 20 else 10)
        return MelHuBERTModel(MelHuBERTConfig(cfg)).to(args.device).eval()
    states = torch.load(args.checkpoint, map_location="cpu", weights_only=False)
    model = MelHuBERTModel(MelHuBERTConfig(states["Upstream_Config"]["melhubert"]))
    if args.mode == "weight-pruning":
        params, _ = get_params_to_prune(model)
        prune.global_unstructured(params, pruning_method=prune.Identity)   # so *_orig / *_mask keys load
        model.load_state_dict(states["model"])
        for module, name in params:
            prune.remove(module, name)                                     # bake the zeros in
        return model.to(args.device)           # NB: the reference leaves dropout on in this mode
    if args.mode == "head-pruning":
        apply_pruned_heads_record(model, states.get("Pruned_heads", []))
        model.load_state_dict(states["model"])
        return model.to(args.device)           # same (Q9)
    model.load_state_dict(states["model"])
    return model.to(args.device).eval()


def main(argv=None):
    args = get_args(argv)
    paths = args.wavs or [os.path.join(EXAMPLE, "100-121669-0000.flac"), os.path.join(EXAMPLE, "1001-134707-0000.flac")]
    if torch.device(args.device).type == "cuda" and not args.host_fbank:
        mel, lens, pad = prepare_data_gpu(paths, args.fp, args.hours, args.device)
    else:
        mel, lens, pad = prepare_data(paths, args.fp, args.hours)
    model = load_model(args)
    with torch.no_grad():
        out = model(mel.to(args.device), pad.to(args.device), get_hidden=True, no_pred=True)
    last_layer_feat, hidden_states = out[0], out[5]
    print(f"[extract_feature] batch {tuple(mel.shape)} lens {lens} -> last layer {tuple(last_layer_feat.shape)}, "
          f"{len(hidden_states)} hidden states")
    return last_layer_feat, hidden_states


if __name__ == "__main__":
    main()
