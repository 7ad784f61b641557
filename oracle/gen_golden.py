"""Generate the committed golden fixtures by EXECUTING THE UNMODIFIED REFERENCE on CPU.

TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).  Usage:

    python oracle/gen_golden.py [section ...]      # default: all sections

Each section writes one small ``tests/golden/<section>.npz``.  Inputs/weights come from the
deterministic generators in ``oracle/melhubert_oracle.py`` (``synth_state_dict``,
``synth_batch``) so that the fixtures only have to hold (sub-sampled) *outputs*.
"""
import copy
import hashlib
import os
import sys
import tempfile
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402
from oracle import melhubert_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
LENS = [750, 712, 655, 601]


def sub(t, st=25, sc=32):
    """Sub-sample a (B, T, C) tensor to keep fixtures small."""
    return t[:, ::st, ::sc].contiguous().numpy().astype(np.float32)


def sha16(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()[:16]


def base_cfg(frame=20, layers=12, drop=0.0):
    cfg = copy.deepcopy(ref_shim.load_yaml("upstream/melhubert/config/config_model.yaml"))["melhubert"]
    if frame == 10:
        cfg.update(feat_emb_dim=40, mask_length=10)
    cfg.update(encoder_layers=layers, dropout=drop, attention_dropout=drop, activation_dropout=drop)
    return cfg


def ref_model(cfg, sd=None):
    from model import MelHuBERTModel, MelHuBERTConfig

    m = MelHuBERTModel(MelHuBERTConfig(cfg))
    if sd is not None:
        m.load_state_dict(sd)
    return m


def save(name, **arrs):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"[golden] {name}: {os.path.getsize(path) / 1024:.1f} KiB, keys={sorted(arrs)}")


# ---------------------------------------------------------------------------------------
def sec_span_mask():
    """Reference compute_mask_indices under np.random.seed(1337) (bit-exact object)."""
    from fairseq_code import compute_mask_indices

    out = {}
    cases = {
        "c20": (4, 750, LENS, 0.7, 5),
        "c10": (4, 1500, [1500, 1420, 1311, 1203], 0.7, 10),
        "short": (3, 40, [40, 17, 9], 0.8, 10),
        "e1": (2, 791, [101, 791], 0.7, 5),
    }
    for key, (b, t, lens, p, ml) in cases.items():
        pad = torch.zeros(b, t, dtype=torch.bool)
        for i, l in enumerate(lens):
            pad[i, l:] = True
        np.random.seed(1337)
        m = compute_mask_indices((b, t), pad, p, ml, "static", 0.0, min_masks=2, no_overlap=False,
                                 min_space=1, require_same_masks=False)
        nxt = np.random.rand()  # pins the amount of RNG consumed
        out[key + "_mask"] = np.packbits(m)
        out[key + "_shape"] = np.array([b, t, p * 1000, ml] + list(lens))
        out[key + "_next"] = np.array([nxt])
        print(key, m.sum(1), sha16(m))
    save("span_mask", **out)


def sec_init():
    """Random-init hashes under seed 1337 -- device independent because
    init_bert_params draws on the CPU generator."""
    import random

    random.seed(1337); np.random.seed(1337); torch.manual_seed(1337)
    cfg = copy.deepcopy(ref_shim.load_yaml("upstream/melhubert/config/config_model.yaml"))["melhubert"]
    m = ref_model(cfg)
    sd = m.state_dict()
    names = sorted(sd)
    hashes = [sha16(sd[n].numpy()) for n in names]
    nparams = sum(p.numel() for p in m.parameters())
    print("params", nparams, hashes[:3])
    save("init_1337", names=np.array(names), hashes=np.array(hashes), nparams=np.array([nparams]),
         fc1_head=sd["encoder.layers.0.fc1.weight"][0, :8].numpy())


def sec_forward20():
    """Base 12L 20 ms: eval forward (no_pred,get_hidden) + masked train forward/backward (dropout 0)."""
    cfg = base_cfg(20, 12)
    sd = O.synth_state_dict(cfg, seed=7)
    m = ref_model(cfg, sd)
    feat, label, pad = O.synth_batch(4, 750, 80, LENS)
    m.eval()
    with torch.no_grad():
        out = m(feat.clone(), pad, get_hidden=True, no_pred=True)
    arrs = {"eval_hidden": sub(out[0]), "eval_pre_feat": sub(out[6])}
    arrs["eval_layers"] = np.stack([sub(h) for h in out[5]])
    arrs["eval_absmean"] = np.array([float(h.abs().mean()) for h in out[5]])
    m.train()
    np.random.seed(1337)
    hidden, logit_m, _, label_m, _, _, _, mask_idx = m(feat.clone(), pad, label, mask=True)
    loss = torch.nn.CrossEntropyLoss(ignore_index=-100, reduction="mean")(logit_m, label_m)
    loss.backward()
    arrs["train_mask"] = np.packbits(mask_idx.numpy())
    arrs["train_label_m"] = label_m.numpy()
    arrs["train_logit_m"] = logit_m.detach()[::7, ::16].numpy()
    arrs["train_loss"] = np.array([float(loss)])
    gn, gs = {}, {}
    for n, p in m.named_parameters():
        gn[n] = float(p.grad.norm())
    keys = sorted(gn)
    arrs["grad_names"] = np.array(keys)
    arrs["grad_norms"] = np.array([gn[k] for k in keys])
    pick = ["encoder.layers.0.fc1.weight", "encoder.layers.11.self_attn.q_proj.weight",
            "encoder.layers.5.self_attn.out_proj.weight", "final_proj.weight", "pre_extract_proj.weight"]
    for k in pick:
        g = dict(m.named_parameters())[k].grad
        arrs["grad_" + k] = g[::37, ::29].numpy()
    arrs["grad_pos_g"] = dict(m.named_parameters())["encoder.pos_conv.0.weight_g"].grad.numpy()
    print("loss", float(loss), "N_m", label_m.numel())
    save("forward20", **arrs)


LENS10 = [1500, 1311]


def sec_forward10():
    """cfg3's shape: base 12L 10 ms (D_in = 40, 1500 frames, mask spans of 10): eval forward + masked train
    forward / backward (dropout 0) on two utterances."""
    cfg = base_cfg(10, 12)
    sd = O.synth_state_dict(cfg, seed=13)
    m = ref_model(cfg, sd)
    feat, label, pad = O.synth_batch(2, 1500, 40, LENS10, seed=21)
    m.eval()
    with torch.no_grad():
        out = m(feat.clone(), pad, get_hidden=True, no_pred=True)
    arrs = {"eval_hidden": sub(out[0], 50, 32), "eval_pre_feat": sub(out[6], 50, 32)}
    arrs["eval_layers"] = np.stack([sub(h, 50, 32) for h in out[5]])
    arrs["eval_absmean"] = np.array([float(h.abs().mean()) for h in out[5]])
    m.train()
    np.random.seed(1337)
    hidden, logit_m, _, label_m, _, _, _, mask_idx = m(feat.clone(), pad, label, mask=True)
    loss = torch.nn.CrossEntropyLoss(ignore_index=-100, reduction="mean")(logit_m, label_m)
    loss.backward()
    arrs["train_mask"] = np.packbits(mask_idx.numpy())
    arrs["train_label_m"] = label_m.numpy()
    arrs["train_logit_m"] = logit_m.detach()[::11, ::16].numpy()
    arrs["train_loss"] = np.array([float(loss)])
    gn = {n: float(p.grad.norm()) for n, p in m.named_parameters()}
    keys = sorted(gn)
    arrs["grad_names"] = np.array(keys)
    arrs["grad_norms"] = np.array([gn[k] for k in keys])
    print("loss", float(loss), "N_m", label_m.numel())
    save("forward10", **arrs)


class _Holder:
    """Stands in for the expert object the tools poke (needs .model / .upstream_config)."""

    def __init__(self, model, cfg):
        self.model = model
        self.upstream_config = {"melhubert": cfg}


def sec_head_prune():
    """Reference HeadPruningTools (l1, by_layer then by_whole) on the 10 ms config."""
    from head_pruning.hp_utils import HeadPruningTools

    cfg = base_cfg(10, 12)
    sd = O.synth_state_dict(cfg, seed=11)
    feat, label, pad = O.synth_batch(2, 300, 40, [300, 233], seed=5)
    arrs = {}
    for target in ("by_layer", "by_whole"):
        m = ref_model(cfg, sd)
        tmp = tempfile.mkdtemp()
        rc = {"prune": {"metric": "l1", "target": target, "total_steps": 11, "num_heads_each_step": 12}}
        tools = HeadPruningTools(Namespace(expdir=tmp, device="cpu"), rc, {"melhubert": cfg}, _Holder(m, cfg))
        scores0 = tools.get_heads_norm(m.encoder)
        arrs[f"{target}_scores0"] = np.array([s for _, s in scores0], dtype=np.float64)
        for step in range(3):
            tools.prune_api()
        # flatten the Pruned_heads record: rows (step, layer, head) in insertion order
        rec = [(s, l, h) for s, grp in enumerate(tools.pruned_heads) for l, hs in grp.items() for h in hs]
        arrs[f"{target}_record"] = np.array(rec)
        arrs[f"{target}_heads"] = np.array([lyr.self_attn.num_heads for lyr in m.encoder.layers])
        m.eval()
        with torch.no_grad():
            out = m(feat.clone(), pad, get_hidden=True, no_pred=True)
        arrs[f"{target}_hidden"] = sub(out[0], 10, 32)
        arrs[f"{target}_qshape"] = np.array(m.encoder.layers[0].self_attn.q_proj.weight.shape)
        print(target, tools.pruned_heads)
    save("head_prune", **arrs)


def sec_row_prune():
    """Reference RowPruningTools: 2 steps of 128 rows on a 4-layer 20 ms model."""
    from row_pruning.rp_utils import RowPruningTools

    cfg = base_cfg(20, 4)
    sd = O.synth_state_dict(cfg, seed=13)
    m = ref_model(cfg, sd)
    rc = {"prune": {"num_rows_each_step": 128, "total_steps": 20}}
    holder = _Holder(m, cfg)
    tools = RowPruningTools(Namespace(expdir=tempfile.mkdtemp(), device="cpu"), rc, {"melhubert": cfg}, holder)
    arrs = {}
    sc = tools.get_layer_rows_norm(m.encoder.layers[0].fc1, m.encoder.layers[0].fc2, 0)
    arrs["scores_l0"] = np.array([s for _, s in sc], dtype=np.float64)
    kept_hash = []
    for step in range(2):
        tools.prune_api()
        kept_hash.append([sha16(l.fc1.bias.detach().numpy()) for l in m.encoder.layers])
    arrs["bias_hash"] = np.array(kept_hash)
    arrs["ffn_dim"] = np.array([tools.total_ffn_dim, holder.upstream_config["melhubert"]["encoder_ffn_embed_dim"]])
    arrs["fc1_shape"] = np.array(m.encoder.layers[3].fc1.weight.shape)
    arrs["fc2_shape"] = np.array(m.encoder.layers[3].fc2.weight.shape)
    feat, label, pad = O.synth_batch(2, 200, 80, [200, 150], seed=6)
    m.eval()
    with torch.no_grad():
        out = m(feat.clone(), pad, get_hidden=True, no_pred=True)
    arrs["hidden"] = sub(out[0], 10, 32)
    save("row_prune", **arrs)


def sec_weight_prune():
    """Reference global_unstructured(L1Unstructured) at 0.5 then (remove +) 0.55 on the base model."""
    from pytorch_code import prune
    from weight_pruning.wp_utils import get_params_to_prune

    cfg = base_cfg(20, 12)
    sd = O.synth_state_dict(cfg, seed=7)
    m = ref_model(cfg, sd)
    params, _ = get_params_to_prune(m)
    prune.global_unstructured(params, pruning_method=prune.Identity)
    arrs = {"keys_identity": np.array(sorted(m.state_dict().keys()))}
    names = O.prunable_names(12)
    for tag, amount in (("s50", 0.5), ("s55", 0.55)):
        for module, name in params:
            prune.remove(module, name)
        prune.global_unstructured(params, pruning_method=prune.L1Unstructured, amount=amount)
        st = m.state_dict()
        counts, hashes = [], []
        for n in names:
            mk = st[n + "_mask"].numpy()
            counts.append(int((~mk).sum()))
            hashes.append(sha16(np.packbits(mk)))
        arrs[tag + "_counts"] = np.array(counts)
        arrs[tag + "_hashes"] = np.array(hashes)
        print(tag, sum(counts))
    arrs["names"] = np.array(names)
    feat, label, pad = O.synth_batch(2, 200, 80, [200, 150], seed=6)
    m.eval()
    with torch.no_grad():
        out = m(feat.clone(), pad, get_hidden=True, no_pred=True)
    arrs["hidden_s55"] = sub(out[0], 10, 32)
    # gradient w.r.t. weight_orig is masked: pin that on one tensor
    m.train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    np.random.seed(1337)
    _, logit_m, _, label_m, *_ = m(feat.clone(), pad, label, mask=True)
    loss = torch.nn.functional.cross_entropy(logit_m, label_m)
    loss.backward()
    g = m.encoder.layers[0].fc1.weight_orig.grad
    mk = m.encoder.layers[0].fc1.weight_mask
    arrs["grad_masked_zero"] = np.array([float(g[~mk].abs().max()), float(g[mk].abs().max())])
    arrs["loss_s55"] = np.array([float(loss)])
    save("weight_prune", **arrs)


def sec_distill():
    """Reference MelHuBERTDistiller (distillation/pretrain_expert.py): 12L teacher -> 2L student."""
    from distillation.pretrain_expert import MelHuBERTDistiller

    tcfg = base_cfg(20, 12)
    scfg = base_cfg(20, 2)
    for c in (tcfg, scfg):
        c.update(skip_masked=False, skip_nomask=False)
    scfg["initial_from_teacher"] = False
    tsd = O.synth_state_dict(tcfg, seed=7)
    ssd = O.synth_state_dict(scfg, seed=21)
    tmp = tempfile.mkdtemp()
    ck = os.path.join(tmp, "teacher.ckpt")
    torch.save({"model": tsd}, ck)
    feat, label, pad = O.synth_batch(4, 750, 80, LENS)
    arrs = {}
    for tag, (ltype, alpha, T) in {"masked": ("masked", 0.5, 2.0), "nomasked": ("nomasked", 1.0, 1.0)}.items():
        ucfg = {"melhubert": dict(scfg), "teacher": dict(tcfg), "loss_param": {"T": T, "alpha": alpha, "type": ltype}}
        ex = MelHuBERTDistiller(ucfg, ck, "cpu", False)
        ex.model.load_state_dict(ssd)
        ex.train()
        np.random.seed(1337)
        loss = ex((feat.clone(), label, pad, LENS))
        loss.backward()
        arrs[tag + "_loss"] = np.array([float(loss)])
        arrs[tag + "_gnorm_fc1"] = np.array([float(ex.model.encoder.layers[1].fc1.weight.grad.norm())])
        arrs[tag + "_gnorm_final"] = np.array([float(ex.model.final_proj.weight.grad.norm())])
        # the individual terms, recomputed through the reference's own loss_fn_kd
        with torch.no_grad():
            np.random.seed(1337)
            t_out = ex.teacher_model(feat.clone(), pad, label, mask=ex.mask_or_not)
            s_out = ex.model(feat.clone(), pad, label, mask=ex.mask_or_not, teacher_mask_indices=t_out[7])
            if ltype == "masked":
                terms = ex.loss_fn_kd(s_out[1], s_out[3], t_out[1], T=T, alpha=alpha)
            else:
                terms = ex.loss_fn_kd(s_out[2], s_out[4], t_out[2], T=T, alpha=alpha)
        arrs[tag + "_terms"] = np.array([float(x) for x in terms])
        print(tag, float(loss), [float(x) for x in terms])
    save("distill", **arrs)


def sec_extract():
    """cfg1: reference extract_feature.py logic on the two example FLACs (CPU), random-init
    weights under seed 1337.  ``torchaudio.load`` has no FLAC backend in this container, so
    the PCM comes from the product's own decoder, which is pinned separately by the
    STREAMINFO MD5 known-answer test."""
    import random
    import torchaudio

    sys.path.insert(0, os.path.dirname(HERE))
    from speech_ssl_compression_b200.frontend.flac import decode_flac

    ex_dir = os.path.join(ref_shim.REFERENCE_ROOT, "example")
    mean_std = np.load(os.path.join(ex_dir, "libri-960-mean-std.npy"))
    mean, std = torch.Tensor(mean_std[0].reshape(-1)), torch.Tensor(mean_std[1].reshape(-1))
    mels, md5s = [], []
    for fn in ("100-121669-0000.flac", "1001-134707-0000.flac"):
        pcm, sr, md5_ok, md5_hex = decode_flac(os.path.join(ex_dir, fn))
        assert md5_ok and sr == 16000
        md5s.append(md5_hex)
        wav = torch.from_numpy(pcm.astype(np.float32) / 32768.0).unsqueeze(0)
        y = torchaudio.compliance.kaldi.fbank(wav * (2 ** 15), num_mel_bins=40, sample_frequency=16000,
                                              window_type="hamming", frame_length=25, frame_shift=10)
        y = (y - mean) / std
        odd, even = y[::2, :], y[1::2, :]
        if odd.shape[0] != even.shape[0]:
            even = torch.cat((even, torch.zeros(1, even.shape[1])), dim=0)
        mels.append(torch.cat((odd, even), dim=1))
    lens = [len(x) for x in mels]
    mel = torch.nn.utils.rnn.pad_sequence(mels, batch_first=True)
    pad = torch.ones(mel.shape[:-1])
    for i, l in enumerate(lens):
        pad[i, l:] = 0
    random.seed(1337); np.random.seed(1337); torch.manual_seed(1337)
    cfg = copy.deepcopy(ref_shim.load_yaml("upstream/melhubert/config/config_model.yaml"))["melhubert"]
    m = ref_model(cfg)
    m.eval()
    with torch.no_grad():
        out = m(mel.clone(), pad, get_hidden=True, no_pred=True)
    save("extract_cfg1", mel=mel.numpy().astype(np.float16), mel_f32_sub=mel[:, ::9, ::7].numpy(),
         lens=np.array(lens), md5=np.array(md5s), hidden=sub(out[0], 7, 16),
         layers=np.stack([sub(h, 25, 32) for h in out[5]]), shape=np.array(out[0].shape))
    print("cfg1", out[0].shape, lens)


SECTIONS = {
    "span_mask": sec_span_mask,
    "init": sec_init,
    "forward20": sec_forward20,
    "forward10": sec_forward10,
    "head_prune": sec_head_prune,
    "row_prune": sec_row_prune,
    "weight_prune": sec_weight_prune,
    "distill": sec_distill,
    "extract": sec_extract,
}

if __name__ == "__main__":
    ref_shim.install()
    torch.set_num_threads(os.cpu_count())
    todo = sys.argv[1:] or list(SECTIONS)
    for s in todo:
        print(f"== {s}")
        SECTIONS[s]()
