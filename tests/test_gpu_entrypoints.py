"""Entry points on the B200: ``extract_feature.py`` (BASELINE.json config 1: the two example FLACs, random init
seed 1337) against the reference golden, and ``train.py -m MODE`` end to end on synthetic buckets with the
reference's yaml schema, prune schedule and checkpoint format."""
import os
import sys

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_extract_feature_cli_matches_reference_golden(golden):
    import extract_feature

    g = golden("extract_cfg1")
    last, hiddens = extract_feature.main(["-m", "melhubert", "-f", "20", "-d", "960"])
    assert tuple(last.shape) == (2, 791, 768) and len(hiddens) == 12
    got = last.float().cpu()[:, ::7, ::16].numpy()
    want = g["hidden"]
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))  # noqa: E731
    assert rel(got[1], want[1]) < 2.5e-2                    # the 791-frame utterance
    n0 = (101 + 6) // 7
    assert rel(got[0, :n0], want[0, :n0]) < 2.5e-2          # valid frames of the 101-frame utterance
    # the front-end itself (FLAC decode + Kaldi fbank + normalisation + 20 ms stacking) is exact
    mel, lens, pad = extract_feature.prepare_data(
        [os.path.join(extract_feature.EXAMPLE, f) for f in ("100-121669-0000.flac", "1001-134707-0000.flac")], 20, 960)
    assert lens == [int(x) for x in g["lens"]]
    np.testing.assert_allclose(mel[:, ::9, ::7].numpy(), g["mel_f32_sub"], rtol=1e-5, atol=1e-5)


def _write_cfgs(tmp, mode, layers=2, accum=1):
    model = {"melhubert": dict(feat_emb_dim=80, encoder_layers=layers, encoder_embed_dim=768, encoder_ffn_embed_dim=3072,
                               encoder_attention_heads=12, num_cluster=512, mask_prob=0.7, mask_length=5, skip_masked=False,
                               skip_nomask=True, dropout=0.1, attention_dropout=0.1, activation_dropout=0.1),
             "task": {"sequence_length": 256}}
    runner = {"runner": {"total_steps": 6, "log_step": 2, "gradient_clipping": 10.0, "save_every_x_epochs": 1,
                         "gradient_accumulate_steps": accum},
              "optimizer": {"lr": 1e-4, "betas": [0.9, 0.999], "eps": 1e-8, "weight_decay": 0},
              "datarc": {"train_batch_size": 2}}
    if mode == "row-pruning":
        runner["prune"] = {"num_rows_each_step": 128, "total_steps": 2, "interval": 2, "warm_up": 2}
    if mode == "head-pruning":
        runner["prune"] = {"metric": "l1", "target": "by_layer", "total_steps": 2, "interval": 2, "warm_up": 2}
    if mode == "weight-pruning":
        runner["prune"] = {"pruning_condition": "fixed", "strategy": "L1Unstructured", "n_iters": 2, "warnup": 2,
                           "period": 2, "sparsity": [0.2, 0.4]}
    mp, rp = os.path.join(tmp, "model.yaml"), os.path.join(tmp, "runner.yaml")
    yaml.safe_dump(model, open(mp, "w"))
    yaml.safe_dump(runner, open(rp, "w"))
    return mp, rp


@pytest.mark.parametrize("mode", ["melhubert", "row-pruning", "head-pruning", "weight-pruning"])
def test_train_cli_runs_each_mode_and_writes_reference_checkpoints(tmp_path, mode):
    import train

    mp, rp = _write_cfgs(str(tmp_path), mode)
    exp = str(tmp_path / "exp")
    train.main(["-m", mode, "-g", mp, "-c", rp, "-n", exp, "-f", "20", "--synthetic"])
    st = torch.load(os.path.join(exp, "last-step.ckpt"), map_location="cpu", weights_only=False)
    assert {"Optimizer", "Step", "Args", "Runner", "model", "Upstream_Config"} <= set(st)
    assert st["Step"] == 6 and os.path.isfile(os.path.join(exp, "config_runner.yaml"))
    sd = st["model"]
    if mode == "row-pruning":
        assert sd["encoder.layers.0.fc1.weight"].shape == (3072 - 256, 768)
        assert st["Upstream_Config"]["melhubert"]["encoder_ffn_embed_dim"] == 3072 - 256
        assert os.path.isfile(os.path.join(exp, "states_prune_3072.ckpt"))
    if mode == "head-pruning":
        assert sd["encoder.layers.0.self_attn.q_proj.weight"].shape == (64 * 10, 768)
        assert len(st["Pruned_heads"]) == 2
    if mode == "weight-pruning":
        m = sd["encoder.layers.0.fc1.weight_mask"]
        assert m.dtype == torch.bool and "encoder.layers.0.fc1.weight_orig" in sd and "Pruning" in st
        total = sum(v.numel() for k, v in sd.items() if k.endswith("_mask"))
        pruned = sum(int((~v).sum()) for k, v in sd.items() if k.endswith("_mask"))
        assert abs(pruned / total - 0.4) < 1e-3
    losses = [float(r.split(",")[2]) for r in open(os.path.join(exp, "train_log.csv")).read().strip().splitlines()]
    assert len(losses) == 3 and all(np.isfinite(losses))
    # "Optimizer" is a torch.optim.Adam state dict over expert.parameters() (runner.py:154-172)
    opt = st["Optimizer"]
    assert set(opt) == {"state", "param_groups"} and float(opt["state"][0]["step"]) > 0


def test_train_cli_gradient_accumulation_and_optimizer_resume(tmp_path):
    """``gradient_accumulate_steps: 2``: 6 optimizer steps = 12 micro-batches (graph-captured pair of graphs); then
    ``-i last-step.ckpt --init_optimizer_from_initial_weight`` restores the Adam moments (runner.py:163-170)."""
    import train

    mp, rp = _write_cfgs(str(tmp_path), "melhubert", accum=2)
    exp = str(tmp_path / "exp")
    train.main(["-m", "melhubert", "-g", mp, "-c", rp, "-n", exp, "-f", "20", "--synthetic"])
    ck = os.path.join(exp, "last-step.ckpt")
    st = torch.load(ck, map_location="cpu", weights_only=False)
    assert st["Step"] == 6 and float(st["Optimizer"]["state"][0]["step"]) == 6
    rows = open(os.path.join(exp, "train_log.csv")).read().strip().splitlines()
    assert [int(r.split(",")[0]) for r in rows] == [2, 4, 6]
    exp2 = str(tmp_path / "exp2")
    train.main(["-m", "melhubert", "-g", mp, "-c", rp, "-n", exp2, "-f", "20", "--synthetic", "-i", ck,
                "--init_optimizer_from_initial_weight", "--max_steps", "2"])
    st2 = torch.load(os.path.join(exp2, "last-step.ckpt"), map_location="cpu", weights_only=False)
    assert float(st2["Optimizer"]["state"][0]["step"]) == 8


def test_train_cli_distillation(tmp_path):
    import train
    from speech_ssl_compression_b200.model import MelHuBERTConfig, MelHuBERTModel

    mp, rp = _write_cfgs(str(tmp_path), "distillation", layers=1)
    cfg = yaml.safe_load(open(mp))
    cfg["teacher"] = dict(cfg["melhubert"], encoder_layers=2, skip_nomask=False)
    cfg["melhubert"]["skip_nomask"] = False
    cfg["loss_param"] = {"T": 1, "alpha": 1, "type": "nomasked"}
    yaml.safe_dump(cfg, open(mp, "w"))
    ck = str(tmp_path / "teacher.ckpt")
    torch.save({"model": MelHuBERTModel(MelHuBERTConfig(cfg["teacher"])).state_dict()}, ck)
    exp = str(tmp_path / "exp")
    train.main(["-m", "distillation", "-g", mp, "-c", rp, "-n", exp, "-i", ck, "--synthetic"])
    st = torch.load(os.path.join(exp, "last-step.ckpt"), map_location="cpu", weights_only=False)
    assert st["Step"] == 6 and not any(k.startswith("teacher") for k in st["model"])


@pytest.mark.parametrize("extra", [[], ["--mode", "extract"], ["--mode", "row+weight", "--accum", "2"]])
def test_bench_prints_one_contract_line(extra):
    """bench.py end to end on a small shape: one JSON line with the keys the measurement contract names."""
    import json
    import subprocess

    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--batch", "2", "--frames", "256",
           "--no-cpu-baseline", "--no-gpu-baseline"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "roofline_attention"):
        assert key in d, key
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    assert 0 < d["roofline"]["frac"] < 1.2 and d["config"]["workload"].split(":")[0] in ("cfg2", "cfg4", "extraction")
