"""Front-end: mel matrix / frame stacking host logic on CPU, ``mh_fbank`` parity on the GPU against
the golden produced by the reference's own front-end (torchaudio.compliance.kaldi.fbank on CPU, oracle/gen_golden.py
section extract_cfg1) and against torchaudio run in the test process."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
EXAMPLE = os.path.join(GOLD, "example")
PATHS = [os.path.join(EXAMPLE, "100-121669-0000.flac"), os.path.join(EXAMPLE, "1001-134707-0000.flac")]


def test_mel_banks_match_torchaudio():
    ta = pytest.importorskip("torchaudio")
    from speech_ssl_compression_b200.frontend.fbank import mel_banks

    ref, _ = ta.compliance.kaldi.get_mel_banks(40, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    ours = mel_banks()
    assert ours.shape == (40, 257) and ours.dtype == np.float32
    assert np.abs(ours[:, :256] - ref.numpy()).max() < 2e-5   # float32 evaluation order differs, nothing else
    assert (ours[:, 256] == 0).all()
    # every filter is a non-empty triangle inside [20 Hz, Nyquist)
    assert (ours.sum(1) > 0).all() and ours.min() >= 0 and ours.max() <= 1.0 + 1e-6


def test_frame_stacking_matches_reference_rule():
    from speech_ssl_compression_b200.frontend.fbank import num_frames, stack_frames

    assert num_frames(32640) == 202 and num_frames(253280) == 1581 and num_frames(399) == 0 and num_frames(400) == 1
    # reference (extract_feature.py:46-52): odd = y[::2], even = y[1::2] (+ a zero row when lengths differ), cat(dim=1)
    for nf in (6, 7):
        y = torch.arange(nf * 3, dtype=torch.float32).reshape(nf, 3) + 1
        odd, even = y[::2], y[1::2]
        if odd.shape[0] != even.shape[0]:
            even = torch.cat((even, torch.zeros(1, 3)), 0)
        ref = torch.cat((odd, even), 1)
        padded = torch.zeros(1, 9, 3)
        padded[0, :nf] = y          # frames past the utterance are exact zeros, like mh_fbank writes them
        got, lens = stack_frames(padded, [nf], 20)
        assert lens == [ref.shape[0]]
        assert torch.equal(got[0, : lens[0]], ref)
    same, lens = stack_frames(padded, [7], 10)
    assert same is padded and lens == [7]


@pytest.mark.gpu
def test_fbank_matches_reference_front_end_golden():
    import extract_feature as EF

    g = np.load(os.path.join(GOLD, "extract_cfg1.npz"))
    mel, lens, pad = EF.prepare_data_gpu(PATHS, 20, 960, "cuda")
    assert list(lens) == [int(x) for x in g["lens"]]
    assert tuple(mel.shape) == tuple(g["mel"].shape)
    assert torch.equal(pad.sum(1).long().cpu(), torch.tensor(lens))
    got = mel.cpu().numpy()
    # normalised log-mel units (std-normalised, O(1)): fp32 FFT / log against torchaudio's, tolerance 2e-3
    assert np.abs(got[:, ::9, ::7] - g["mel_f32_sub"]).max() < 2e-3
    # the full tensor is stored as fp16 in the fixture (|x| < 16 -> spacing <= 2^-7)
    assert np.abs(got - g["mel"].astype(np.float32)).max() < 1e-2
    # frames past each utterance's end are exact zeros (the pad mask convention of the encoder)
    assert float(mel[0, lens[0]:].abs().max()) == 0.0


@pytest.mark.gpu
def test_fbank_vs_torchaudio_ragged_batch():
    ta = pytest.importorskip("torchaudio")
    from speech_ssl_compression_b200.frontend.fbank import kaldi_fbank

    g = torch.Generator().manual_seed(11)
    lens = [16000, 9999, 400, 4801]
    waves = [(torch.randn(n, generator=g) * 0.1 + 0.02).clamp(-1, 1) for n in lens]
    feat, frames = kaldi_fbank(waves, device="cuda")
    assert frames == [98, 60, 1, 28]
    for i, w in enumerate(waves):
        ref = ta.compliance.kaldi.fbank(w[None] * 2 ** 15, num_mel_bins=40, sample_frequency=16000, window_type="hamming",
                                        frame_length=25, frame_shift=10)
        assert ref.shape[0] == frames[i]
        assert (feat[i, : frames[i]].cpu() - ref).abs().max() < 1e-3   # log-mel, natural log units
        assert float(feat[i, frames[i]:].abs().max()) == 0.0 if frames[i] < feat.shape[1] else True


@pytest.mark.gpu
def test_s3prl_upstream_expert_end_to_end_matches_reference_golden(tmp_path):
    """Waveforms -> mh_fbank -> encoder through the S3PRL-style expert (s3prl_upstream/expert.py:113-139) against the
    golden of the reference's extract_feature.py on the same FLACs and the same random-init (seed 1337) weights."""
    import random

    import extract_feature as EF
    from speech_ssl_compression_b200.model import MelHuBERTConfig, MelHuBERTModel
    from speech_ssl_compression_b200.s3prl_upstream import UpstreamExpert

    g = np.load(os.path.join(GOLD, "extract_cfg1.npz"))
    random.seed(1337); np.random.seed(1337); torch.manual_seed(1337)
    cfg = dict(feat_emb_dim=80, encoder_layers=12, mask_prob=0.7, mask_length=5)
    m = MelHuBERTModel(MelHuBERTConfig(cfg))
    ckpt = str(tmp_path / "random_init.ckpt")
    torch.save({"Upstream_Config": {"melhubert": cfg}, "model": m.state_dict()}, ckpt)
    ex = UpstreamExpert(ckpt, mode="melhubert", fp=20, mean_std_npy_path=os.path.join(EXAMPLE, "libri-960-mean-std.npy"))
    ex = ex.to("cuda").eval()
    assert ex.get_downsample_rates("hidden_states") == 320
    wavs = [EF.load_waveform(p)[0].reshape(-1).to("cuda") for p in PATHS]
    with torch.no_grad():
        st = ex(wavs)
    assert set(st) == {"hidden_states", "last_hidden_state"} and len(st["hidden_states"]) == 13
    assert tuple(st["last_hidden_state"].shape) == tuple(int(x) for x in g["shape"])
    got = st["last_hidden_state"][:, ::7, ::16].float().cpu().numpy()
    want = g["hidden"]
    rel = np.linalg.norm(got[1] - want[1]) / np.linalg.norm(want[1])
    assert rel < 2.5e-2, rel      # bf16 pipeline vs the fp32 reference
    with pytest.raises(RuntimeError):
        ex([w.cpu() for w in wavs])
