"""CPU-only checks: the C-ABI library loads and exports every symbol include/mh_b200.h declares (no
compute calls without a GPU), and the host-side mirror of the reference (config defaults, parameter
names + random init, span masks, prune selections, checkpoint surgery) matches the reference goldens."""
import ctypes
import hashlib
import os
import re
import subprocess
import tempfile
from argparse import Namespace

import numpy as np
import pytest
import torch

from oracle import melhubert_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "speech_ssl_compression_b200")


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# ---------------------------------------------------------------------------------------------- ABI
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mh_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mh_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    so = os.path.join(PKG, "libmh_b200.so")
    if not os.path.isfile(so):
        subprocess.run(["make", "-C", os.path.join(PKG, "csrc"), "-j", str(os.cpu_count() or 4)], check=True)
    lib = ctypes.CDLL(so)
    names = _declared_symbols()
    assert len(names) >= 40, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.mh_version() >= 100
    lib.mh_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.mh_last_error(), bytes)


def test_argument_validation_fails_loudly_before_any_device_work():
    """Shape / alignment violations are errors with a message (never a fallback, never a device fault): the checks
    sit in front of the first CUDA call, so they can be exercised without a GPU."""
    lib = ctypes.CDLL(os.path.join(PKG, "libmh_b200.so"))
    lib.mh_last_error.restype = ctypes.c_char_p
    vp, fp = ctypes.c_void_p, ctypes.c_void_p
    buf = ctypes.create_string_buffer(4096)
    base = (ctypes.addressof(buf) + 63) & ~63          # 64-byte aligned host address (never dereferenced)
    lib.mh_layernorm_fwd.argtypes = [vp, fp, fp, vp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                     ctypes.c_uint64, ctypes.c_uint32, vp]
    lib.mh_colsum.argtypes = [vp, ctypes.c_longlong, fp, ctypes.c_int, ctypes.c_int, vp]
    # columns not a multiple of 8
    assert lib.mh_layernorm_fwd(base, base, base, base, base, base, 4, 60, 1e-5, 0.0, 0, 0, None) != 0
    assert b"bad shape" in lib.mh_last_error()
    # misaligned row pointer
    assert lib.mh_layernorm_fwd(base + 2, base, base, base, base, base, 4, 64, 1e-5, 0.0, 0, 0, None) != 0
    assert b"16-byte aligned" in lib.mh_last_error()
    # null output
    assert lib.mh_layernorm_fwd(base, base, base, None, base, base, 4, 64, 1e-5, 0.0, 0, 0, None) != 0
    assert b"null" in lib.mh_last_error()
    # column sums: leading dimension smaller than the width, misaligned base
    assert lib.mh_colsum(base, 8, base, 4, 64, None) != 0
    assert b"bad shape" in lib.mh_last_error()
    assert lib.mh_colsum(base + 2, 64, base, 4, 64, None) != 0
    assert b"aligned" in lib.mh_last_error()


def test_header_cites_the_reference_for_every_kernel_family():
    text = open(os.path.join(ROOT, "include", "mh_b200.h")).read()
    for cite in ("forward_multihead_attention.py", "module.py", "model.py", "pretrain_expert.py", "prune.py", "runner.py"):
        assert cite in text, cite


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch it --
    nothing under the product package, the entry points or tools/ may."""
    paths = []
    for top in (PKG, os.path.join(ROOT, "tools")):
        for dirpath, _, files in os.walk(top):
            paths += [os.path.join(dirpath, f) for f in files if f.endswith(".py")]
    paths += [os.path.join(ROOT, f) for f in ("train.py", "extract_feature.py", "runner.py")]
    assert len(paths) > 20
    for path in paths:
        src = open(path).read()
        assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), path


def test_model_refuses_cpu_tensors():
    from speech_ssl_compression_b200.model import MelHuBERTConfig, MelHuBERTModel

    m = MelHuBERTModel(MelHuBERTConfig(dict(feat_emb_dim=80, encoder_layers=1)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 16, 80), torch.ones(1, 16))


# ------------------------------------------------------------------------------- config / init parity
def test_config_defaults_match_reference():
    from speech_ssl_compression_b200.model_config import MelHuBERTConfig

    c = MelHuBERTConfig({})
    assert (c.feat_emb_dim, c.encoder_layers, c.encoder_embed_dim, c.encoder_ffn_embed_dim, c.encoder_attention_heads) == (40, 1, 768, 3072, 12)
    assert (c.mask_prob, c.mask_length, c.num_cluster, c.skip_nomask, c.skip_masked) == (0.8, 10, 512, True, False)
    assert (c.dropout, c.attention_dropout, c.activation_dropout, c.encoder_layerdrop) == (0.1, 0.1, 0.1, 0.0)
    assert (c.conv_pos, c.conv_pos_groups, c.layer_norm_first, c.mask_before_proj) == (128, 16, False, True)


def test_random_init_is_bit_identical_to_the_reference(golden):
    """Same construction order + init_bert_params on the CPU generator -> the seed-1337 weights of the
    reference model (fixture init_1337.npz: parameter names and sha256 of the raw fp32 bytes)."""
    import random

    from speech_ssl_compression_b200.model import MelHuBERTConfig, MelHuBERTModel

    g = golden("init_1337")
    random.seed(1337); np.random.seed(1337); torch.manual_seed(1337)
    m = MelHuBERTModel(MelHuBERTConfig(dict(feat_emb_dim=80, encoder_layers=12, mask_prob=0.7, mask_length=5)))
    sd = m.state_dict()
    names = [str(n) for n in g["names"]]
    assert sorted(sd.keys()) == sorted(names)                                # state_dict key set
    assert sum(p.numel() for p in m.parameters()) == int(g["nparams"][0])
    got = [sha16(sd[n].numpy()) for n in names]
    want = [str(h) for h in g["hashes"]]
    bad = [n for n, a, b in zip(names, got, want) if a != b]
    assert not bad, bad[:5]


# ------------------------------------------------------------------------------------- span masks
@pytest.mark.parametrize("case", ["c20", "c10", "short", "e1"])
def test_span_mask_host_generator_bit_exact(golden, case):
    from speech_ssl_compression_b200.fairseq_code import compute_mask_indices

    g = golden("span_mask")
    shp = g[case + "_shape"]
    b, t, p, ml = int(shp[0]), int(shp[1]), shp[2] / 1000.0, int(shp[3])
    lens = [int(x) for x in shp[4:]]
    np.random.seed(1337)
    m = compute_mask_indices((b, t), None, p, ml, "static", 0.0, min_masks=2, no_overlap=False, min_space=1,
                             require_same_masks=False, valid_lens=lens)
    nxt = np.random.rand()
    assert np.array_equal(np.packbits(m), g[case + "_mask"])
    assert nxt == g[case + "_next"][0]                                      # same RNG consumption
    # the padding-mask door (what the reference call site passes) gives the same mask
    pm = torch.zeros(b, t, dtype=torch.bool)
    for i, l in enumerate(lens):
        pm[i, l:] = True
    np.random.seed(1337)
    m2 = compute_mask_indices((b, t), pm, p, ml, "static", 0.0, min_masks=2, require_same_masks=False)
    assert np.array_equal(m, m2)


# ------------------------------------------------------------------ pruning objects on CPU tensors
class _Holder:
    def __init__(self, model, cfg):
        self.model, self.upstream_config, self.pruned_heads = model, {"melhubert": cfg}, None


def _cfg(frame, layers):
    return dict(feat_emb_dim=80 if frame == 20 else 40, encoder_layers=layers, mask_prob=0.7,
                mask_length=5 if frame == 20 else 10)


@pytest.mark.parametrize("target", ["by_layer", "by_whole"])
def test_head_pruning_tools_selection_bit_exact_on_cpu(golden, target):
    from speech_ssl_compression_b200.head_pruning.hp_utils import HeadPruningTools
    from speech_ssl_compression_b200.model import MelHuBERTConfig, MelHuBERTModel

    g = golden("head_prune")
    cfg = _cfg(10, 12)
    m = MelHuBERTModel(MelHuBERTConfig(cfg))
    m.load_state_dict(O.synth_state_dict(cfg, seed=11))
    rc = {"prune": {"metric": "l1", "target": target, "total_steps": 11, "num_heads_each_step": 12}}
    tools = HeadPruningTools(Namespace(expdir=tempfile.mkdtemp(), device="cpu"), rc, {"melhubert": cfg}, _Holder(m, cfg))
    s0 = np.array([s for _, s in tools.get_heads_norm(m.encoder)])
    np.testing.assert_allclose(s0, g[f"{target}_scores0"], rtol=1e-6)
    for _ in range(3):
        tools.prune_api()
    rec = [(s, l, h) for s, grp in enumerate(tools.pruned_heads) for l, hs in grp.items() for h in hs]
    assert np.array_equal(np.array(rec), g[f"{target}_record"])
    assert [l.self_attn.num_heads for l in m.encoder.layers] == g[f"{target}_heads"].tolist()
    # sliced tensors equal the oracle's slicing of the same state dict
    sd = O.synth_state_dict(cfg, seed=11)
    for step_grp in tools.pruned_heads:
        for l, hs in step_grp.items():
            O.slice_heads(sd, l, hs)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_row_pruning_tools_selection_bit_exact_on_cpu(golden):
    from speech_ssl_compression_b200.model import MelHuBERTConfig, MelHuBERTModel
    from speech_ssl_compression_b200.row_pruning.rp_utils import RowPruningTools

    g = golden("row_prune")
    cfg = _cfg(20, 4)
    m = MelHuBERTModel(MelHuBERTConfig(cfg))
    m.load_state_dict(O.synth_state_dict(cfg, seed=13))
    holder = _Holder(m, cfg)
    rc = {"prune": {"num_rows_each_step": 128, "total_steps": 20}}
    tools = RowPruningTools(Namespace(expdir=tempfile.mkdtemp(), device="cpu"), rc, {"melhubert": cfg}, holder)
    for step in range(2):
        tools.prune_api()
        assert [sha16(l.fc1.bias.detach().numpy()) for l in m.encoder.layers] == [str(x) for x in g["bias_hash"][step]]
    assert holder.upstream_config["melhubert"]["encoder_ffn_embed_dim"] == 3072 - 256 == m.encoder.ffn_embedding_dim


def test_weight_pruning_reparam_contract_on_cpu(golden):
    """*_orig / *_mask key set, bool masks, Identity -> L1 -> remove life cycle (prune.py:190-296)."""
    from speech_ssl_compression_b200.model import MelHuBERTConfig, MelHuBERTModel
    from speech_ssl_compression_b200.pytorch_code import prune
    from speech_ssl_compression_b200.weight_pruning.wp_utils import get_params_to_prune

    g = golden("weight_prune")
    cfg = _cfg(20, 12)
    m = MelHuBERTModel(MelHuBERTConfig(cfg))
    params, is_prunable = get_params_to_prune(m)
    assert len(params) == 144 and is_prunable("encoder.layers.3.fc1.weight") and not is_prunable("final_proj.weight")
    prune.global_unstructured(params, pruning_method=prune.Identity)
    assert sorted(m.state_dict().keys()) == [str(x) for x in g["keys_identity"]]
    assert prune.is_pruned(m)
    fc1 = m.encoder.layers[0].fc1
    w_param = fc1.weight_orig
    assert fc1.weight_mask.dtype == torch.bool and bool(fc1.weight_mask.all())
    small = [(fc1, "weight"), (fc1, "bias")]
    for mod, name in small:
        prune.remove(mod, name)
    assert fc1.weight is w_param                                             # same Parameter object survives remove()
    prune.global_unstructured(small, pruning_method=prune.L1Unstructured, amount=0.5)
    total = fc1.weight_orig.numel() + fc1.bias_orig.numel()
    pruned = int((~fc1.weight_mask).sum()) + int((~fc1.bias_mask).sum())
    assert pruned == int(round(0.5 * total))
    ref_masks, k, thr, ties = O.global_l1_masks([fc1.weight_orig.detach(), fc1.bias_orig.detach()], 0.5)
    if ties <= 1:
        assert torch.equal(fc1.weight_mask, ref_masks[0]) and torch.equal(fc1.bias_mask, ref_masks[1])
    assert torch.equal(fc1.weight, fc1.weight_orig.masked_fill(~fc1.weight_mask, 0))


def test_flac_decoder_known_answer():
    """STREAMINFO MD5 of the two example files == MD5 of the decoded PCM."""
    from speech_ssl_compression_b200.frontend.flac import decode_flac

    want = {"100-121669-0000.flac": ("1f9b4b53e3c194f950f62649f7147a05", 32640),
            "1001-134707-0000.flac": ("903a2666cf9765de9bb270fdd3fe1f08", 253280)}
    for fn, (md5, n) in want.items():
        pcm, sr, ok, hexd = decode_flac(os.path.join(ROOT, "tests", "golden", "example", fn))
        assert ok and hexd == md5 and sr == 16000 and len(pcm) == n


# ------------------------------------------------------------------------------------ training-loop host logic
def test_fused_adam_state_dict_is_the_torch_adam_layout():
    """The checkpoint's ``"Optimizer"`` entry (runner.py:154-172, mh_utils.py:17) is ``torch.optim.Adam.state_dict()``
    over ``expert.parameters()``: same keys, indices and shapes as torch writes, frozen parameters hold an index but
    no state, and a torch-written state dict loads into the flat moment buffers (resume, runner.py:163-170)."""
    from speech_ssl_compression_b200.parallel import FlatBuffers
    from speech_ssl_compression_b200.trainer import FusedAdam

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    for p in net[1].parameters():
        p.requires_grad_(False)  # (the distillation teacher)
    ref = torch.optim.Adam(net.parameters(), lr=3e-4, betas=(0.8, 0.95), eps=1e-7)
    for _ in range(3):
        net(torch.randn(4, 6)).sum().backward()
        ref.step()
        ref.zero_grad()
    want = ref.state_dict()
    train = [p for p in net.parameters() if p.requires_grad]
    opt = FusedAdam(FlatBuffers(train[::-1]), lr=1.0)  # flat order differs from parameters() order on purpose
    opt.param_order = list(net.parameters())
    opt.load_state_dict(want)
    assert (opt.lr, opt.betas, opt.eps) == (3e-4, (0.8, 0.95), 1e-7) and int(opt.step_count) == 3
    got = opt.state_dict()
    assert set(got) == {"state", "param_groups"} and sorted(got["state"]) == sorted(want["state"]) == [0, 1, 4, 5]
    assert got["param_groups"][0]["params"] == want["param_groups"][0]["params"] == list(range(6))
    for i, st in want["state"].items():
        assert set(got["state"][i]) == {"step", "exp_avg", "exp_avg_sq"} and float(got["state"][i]["step"]) == 3
        assert torch.equal(got["state"][i]["exp_avg"], st["exp_avg"])
        assert torch.equal(got["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
    again = torch.optim.Adam(net.parameters())
    again.load_state_dict(got)  # and torch accepts what we write
    with pytest.raises(ValueError):
        opt.param_order = opt.param_order[:-1]
        opt.load_state_dict(want)


def test_bucket_dataset_is_sharded_by_rank(tmp_path):
    """Data parallelism must add batch, not repeat it: ranks walk disjoint, equally sized shards of one bucket
    permutation (the reference's DataParallel scatters one batch over the devices, pretrain_expert.py:28-30)."""
    import pandas as pd

    import runner

    rows = []
    for i in range(22):
        n = 80 + 3 * i  # >= 40 stacked 20 ms frames: every utterance is cropped to sequence_length
        np.save(tmp_path / f"f{i}.npy", np.full((n, 40), float(i), dtype=np.float32))
        np.save(tmp_path / f"l{i}.npy", np.arange(n) % 7)
        rows.append((str(tmp_path / f"f{i}.npy"), str(tmp_path / f"l{i}.npy"), n))
    pd.DataFrame(rows, columns=["file_path", "label_path", "length"]).to_csv(tmp_path / "set.csv", index=False)
    datarc, task = {"sets": [str(tmp_path / "set.csv")]}, {"sequence_length": 30}
    seen = []
    for rank in range(2):
        torch.manual_seed(1337)
        ds = runner.CsvNpyBuckets(datarc, task, 20, 2, rank=rank, world=2)
        assert len(ds) == 5  # 11 buckets -> 5 per rank, the odd one is dropped so the ranks stay in step
        ids = []
        for feat, label, pad, lens in ds:
            assert feat.shape == (2, 30, 80) and label.shape == (2, 30) and lens == [30, 30]
            ids.append(tuple(sorted({int(feat[j, 0, 0]) for j in range(2)})))
        seen.append(ids)
    assert len(seen[0]) == len(seen[1]) == 5 and not set(seen[0]) & set(seen[1])
