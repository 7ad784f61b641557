"""Log-mel front-end on the GPU.

Host mirror of the feature extraction the reference performs on the CPU in ``extract_feature.py:32-53`` /
``s3prl_upstream/expert.py:23-43``::

    y = torchaudio.compliance.kaldi.fbank(wav * 2**15, num_mel_bins=40, sample_frequency=16000,
                                          window_type='hamming', frame_length=25, frame_shift=10)
    y = (y - mean) / std;  20 ms: y = cat(y[0::2], y[1::2], dim=-1)   (odd last frame dropped)

``torchaudio`` (2.11, un-pinned by the reference) is a third-party dependency; the algorithm restated here is its
published Kaldi-compatible fbank at the defaults the reference leaves untouched (dither 0, preemphasis 0.97,
remove_dc_offset, snip_edges, round_to_power_of_two -> 512-point FFT, use_power, use_log_fbank, low_freq 20,
high_freq 0 -> Nyquist, no VTLN, no energy).  The mel matrix is built here on the host (numpy, float32 like
``get_mel_banks``), everything per-sample runs in ``mh_fbank`` (csrc/frontend.cu).
"""
import functools
import math

import numpy as np
import torch

from .. import kernels as K

WINDOWS = {"hamming": 0, "hanning": 1, "povey": 2, "rectangular": 3}


def _mel(f):
    return 1127.0 * np.log(1.0 + f / 700.0)


@functools.lru_cache(maxsize=8)
def mel_banks(num_bins=40, n_fft=512, sample_freq=16000.0, low_freq=20.0, high_freq=0.0):
    """torchaudio.compliance.kaldi.get_mel_banks (no VTLN) padded with the zero Nyquist column: f32 [num_bins, n_fft/2+1]."""
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    if not (0.0 <= low_freq < nyquist and 0.0 < high_freq <= nyquist and low_freq < high_freq):
        raise ValueError(f"bad mel range: low {low_freq} high {high_freq} nyquist {nyquist}")
    n_bins_fft = n_fft // 2
    bin_width = np.float32(sample_freq / n_fft)
    mel_lo, mel_hi = np.float32(_mel(low_freq)), np.float32(_mel(high_freq))
    delta = np.float32((mel_hi - mel_lo) / (num_bins + 1))
    b = np.arange(num_bins, dtype=np.float32)[:, None]
    left, center, right = mel_lo + b * delta, mel_lo + (b + 1.0) * delta, mel_lo + (b + 2.0) * delta
    mel = _mel(bin_width * np.arange(n_bins_fft, dtype=np.float32)).astype(np.float32)[None, :]
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    w = np.maximum(0.0, np.minimum(up, down)).astype(np.float32)
    return np.concatenate([w, np.zeros((num_bins, 1), np.float32)], axis=1)


def num_frames(n_samples, frame_len=400, frame_shift=160):
    return 1 + (n_samples - frame_len) // frame_shift if n_samples >= frame_len else 0


def kaldi_fbank(waves, mean=None, std=None, *, num_mel_bins=40, sample_frequency=16000.0, frame_length=25.0,
                frame_shift=10.0, window_type="hamming", preemphasis=0.97, scale=2.0 ** 15, device="cuda"):
    """waves: list of 1-D float waveforms in [-1, 1) (any device).  Returns (feat f32 [B, max_frames, num_mel_bins] on
    ``device``, frames list).  ``mean`` / ``std``: optional (num_mel_bins,) normalisation (extract_feature.py:42-44)."""
    flen = int(sample_frequency * frame_length * 0.001)
    fshift = int(sample_frequency * frame_shift * 0.001)
    n_fft = 1 << (flen - 1).bit_length()
    if n_fft != 512:
        raise ValueError(f"mh_fbank is built for a 512-point FFT (frame of {flen} samples needs {n_fft})")
    lens = [int(w.numel()) for w in waves]
    if min(lens) < flen:
        raise ValueError(f"waveform shorter than one frame ({min(lens)} < {flen} samples)")
    L_ = max(lens)
    batch = torch.zeros(len(waves), L_, dtype=torch.float32)
    for i, w in enumerate(waves):
        batch[i, : lens[i]] = w.detach().reshape(-1).float().cpu()
    dev = torch.device(device)
    batch = batch.pin_memory().to(dev, non_blocking=True)
    n = torch.tensor(lens, dtype=torch.int32).to(dev)
    mw = torch.from_numpy(mel_banks(num_mel_bins, n_fft, float(sample_frequency))).to(dev)
    m = s = None
    if mean is not None:
        m = torch.as_tensor(np.asarray(mean), dtype=torch.float32).to(dev).contiguous()
        s = (1.0 / torch.as_tensor(np.asarray(std), dtype=torch.float64)).float().to(dev).contiguous()
    feat = K.fbank(batch, n, mw, mean=m, inv_std=s, frame_len=flen, frame_shift=fshift, scale=scale, preemph=preemphasis,
                   window_type=WINDOWS[window_type])
    return feat, [num_frames(l, flen, fshift) for l in lens]


def stack_frames(feat, frames, fp=20):
    """20 ms frame period: frame pairs (2i, 2i+1) concatenated; an odd last frame is paired with zeros
    (extract_feature.py:46-52).  On the padded batch (frames past an utterance's end are exact zeros) this is
    a reshape.  Returns (feat [B, T', D'], lens)."""
    if fp == 10:
        return feat, list(frames)
    B, F, D = feat.shape
    if F % 2:
        feat = torch.cat([feat, feat.new_zeros(B, 1, D)], dim=1)
        F += 1
    return feat.reshape(B, F // 2, 2 * D), [(f + 1) // 2 for f in frames]
