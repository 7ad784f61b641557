"""Pins the CPU oracle (oracle/melhubert_oracle.py) against fixtures produced by executing the
unmodified reference (oracle/gen_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import melhubert_oracle as O

LENS = [750, 712, 655, 601]


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def sub(t, st=25, sc=32):
    return t[:, ::st, ::sc].contiguous().numpy()


def base_cfg(frame=20, layers=12):
    cfg = dict(feat_emb_dim=80 if frame == 20 else 40, encoder_layers=layers, encoder_embed_dim=768,
               encoder_ffn_embed_dim=3072, encoder_attention_heads=12, num_cluster=512,
               mask_prob=0.7, mask_length=5 if frame == 20 else 10, skip_masked=False, skip_nomask=True)
    return cfg


@pytest.mark.parametrize("case", ["c20", "c10", "short", "e1"])
def test_span_mask_bit_exact(golden, case):
    g = golden("span_mask")
    shp = g[case + "_shape"]
    b, t, p, ml = int(shp[0]), int(shp[1]), shp[2] / 1000.0, int(shp[3])
    lens = [int(x) for x in shp[4:]]
    np.random.seed(1337)
    m = O.span_mask(b, t, lens, p, ml)
    nxt = np.random.rand()
    assert np.array_equal(np.packbits(m), g[case + "_mask"])
    assert nxt == g[case + "_next"][0]  # same amount of RNG consumed


def test_span_mask_survey_known_answer():
    np.random.seed(1337)
    m = O.span_mask(4, 750, LENS, 0.7, 5)
    assert m.sum(1).tolist() == [381, 406, 342, 325]
    assert sha16(m) == "b2dc1ac7c8bfff3e"


@pytest.fixture(scope="module")
def fwd20():
    cfg = base_cfg(20, 12)
    sd = O.synth_state_dict(cfg, seed=7)
    feat, label, pad = O.synth_batch(4, 750, 80, LENS)
    return cfg, sd, feat, label, pad


def test_eval_forward_matches_reference(golden, fwd20):
    cfg, sd, feat, label, pad = fwd20
    g = golden("forward20")
    with torch.no_grad():
        out = O.model_forward(sd, cfg, feat, pad, no_pred=True)
    np.testing.assert_allclose(sub(out["hidden"]), g["eval_hidden"], atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(sub(out["pre_feat"]), g["eval_pre_feat"], atol=1e-5, rtol=1e-5)
    for i, h in enumerate(out["layer_hiddens"]):
        np.testing.assert_allclose(sub(h), g["eval_layers"][i], atol=2e-4, rtol=1e-4)
        assert abs(float(h.abs().mean()) - g["eval_absmean"][i]) < 1e-4
    assert float(out["pre_feat"][3, 700].abs().max()) == 0.0  # padded rows of pre_feat are exactly 0


def test_train_forward_backward_matches_reference(golden, fwd20):
    cfg, sd, feat, label, pad = fwd20
    g = golden("forward20")
    np.random.seed(1337)
    mask = torch.from_numpy(O.span_mask(4, 750, LENS, 0.7, 5))
    assert np.array_equal(np.packbits(mask.numpy()), g["train_mask"])
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = O.model_forward(sdg, cfg, feat, pad, label, mask_indices=mask)
    assert np.array_equal(out["label_m"].numpy(), g["train_label_m"])  # bit-exact label gather
    np.testing.assert_allclose(out["logit_m"].detach()[::7, ::16].numpy(), g["train_logit_m"], atol=3e-4, rtol=1e-4)
    loss = O.ce_mean(out["logit_m"], out["label_m"])
    assert abs(float(loss) - g["train_loss"][0]) < 2e-5
    loss.backward()
    names = [str(n) for n in g["grad_names"]]
    for n, ref in zip(names, g["grad_norms"]):
        got = float(sdg[n].grad.norm())
        assert abs(got - ref) <= 2e-4 * max(ref, 1e-3) + 1e-7, (n, got, ref)
    for n in ["encoder.layers.0.fc1.weight", "encoder.layers.11.self_attn.q_proj.weight", "final_proj.weight"]:
        np.testing.assert_allclose(sdg[n].grad[::37, ::29].numpy(), g["grad_" + n], atol=2e-6, rtol=2e-3)
    np.testing.assert_allclose(sdg["encoder.pos_conv.0.weight_g"].grad.numpy(), g["grad_pos_g"], atol=1e-6, rtol=2e-3)


def test_10ms_forward_backward_matches_reference(golden):
    """cfg3's shape (10 ms frames: D_in = 40, 1500 frames, mask spans of 10) against the reference's own run."""
    lens = [1500, 1311]
    cfg = base_cfg(10, 12)
    sd = O.synth_state_dict(cfg, seed=13)
    feat, label, pad = O.synth_batch(2, 1500, 40, lens, seed=21)
    g = golden("forward10")
    with torch.no_grad():
        out = O.model_forward(sd, cfg, feat, pad, no_pred=True)
    np.testing.assert_allclose(sub(out["hidden"], 50, 32), g["eval_hidden"], atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(sub(out["pre_feat"], 50, 32), g["eval_pre_feat"], atol=1e-5, rtol=1e-5)
    for i, h in enumerate(out["layer_hiddens"]):
        np.testing.assert_allclose(sub(h, 50, 32), g["eval_layers"][i], atol=2e-4, rtol=1e-4)
        assert abs(float(h.abs().mean()) - g["eval_absmean"][i]) < 1e-4
    np.random.seed(1337)
    mask = torch.from_numpy(O.span_mask(2, 1500, lens, 0.7, 10))
    assert np.array_equal(np.packbits(mask.numpy()), g["train_mask"])                     # bit-exact span mask
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = O.model_forward(sdg, cfg, feat, pad, label, mask_indices=mask)
    assert np.array_equal(out["label_m"].numpy(), g["train_label_m"])                     # bit-exact label gather
    np.testing.assert_allclose(out["logit_m"].detach()[::11, ::16].numpy(), g["train_logit_m"], atol=3e-4, rtol=1e-4)
    loss = O.ce_mean(out["logit_m"], out["label_m"])
    assert abs(float(loss) - g["train_loss"][0]) < 2e-5
    loss.backward()
    for n, ref in zip([str(n) for n in g["grad_names"]], g["grad_norms"]):
        got = float(sdg[n].grad.norm())
        assert abs(got - ref) <= 2e-4 * max(ref, 1e-3) + 1e-7, (n, got, ref)


def test_head_scores_and_selection(golden):
    g = golden("head_prune")
    cfg = base_cfg(10, 12)
    for target in ("by_layer", "by_whole"):
        sd = O.synth_state_dict(cfg, seed=11)
        rec = []
        for step in range(3):
            scores = [O.head_scores(sd, l) for l in range(12)]
            if step == 0:
                flat = np.array([s for row in scores for s in row])
                assert np.array_equal(flat, g[f"{target}_scores0"])  # bit-exact doubles
            grp = O.select_heads(scores, target, 12)
            for l, hs in grp.items():
                rec += [(step, l, h) for h in hs]
                O.slice_heads(sd, l, hs)
        assert np.array_equal(np.array(rec), g[f"{target}_record"])
        heads = [sd[f"encoder.layers.{l}.self_attn.q_proj.weight"].shape[0] // 64 for l in range(12)]
        assert heads == g[f"{target}_heads"].tolist()
        feat, label, pad = O.synth_batch(2, 300, 40, [300, 233], seed=5)
        with torch.no_grad():
            out = O.model_forward(sd, cfg, feat, pad, no_pred=True)
        np.testing.assert_allclose(sub(out["hidden"], 10, 32), g[f"{target}_hidden"], atol=2e-4, rtol=1e-4)


def test_row_scores_and_selection(golden):
    g = golden("row_prune")
    cfg = base_cfg(20, 4)
    sd = O.synth_state_dict(cfg, seed=13)
    assert np.array_equal(np.array(O.row_scores(sd, 0)), g["scores_l0"])
    for step in range(2):
        for l in range(4):
            O.slice_rows(sd, l, O.select_rows(O.row_scores(sd, l), 128))
        hashes = [sha16(sd[f"encoder.layers.{l}.fc1.bias"].numpy()) for l in range(4)]
        assert hashes == [str(x) for x in g["bias_hash"][step]]
    assert list(sd["encoder.layers.3.fc1.weight"].shape) == g["fc1_shape"].tolist()
    assert list(sd["encoder.layers.3.fc2.weight"].shape) == g["fc2_shape"].tolist()
    feat, label, pad = O.synth_batch(2, 200, 80, [200, 150], seed=6)
    with torch.no_grad():
        out = O.model_forward(sd, cfg, feat, pad, no_pred=True)
    np.testing.assert_allclose(sub(out["hidden"], 10, 32), g["hidden"], atol=2e-4, rtol=1e-4)


def test_global_l1_masks(golden):
    g = golden("weight_prune")
    cfg = base_cfg(20, 12)
    sd = O.synth_state_dict(cfg, seed=7)
    names = O.prunable_names(12)
    assert names == [str(n) for n in g["names"]]
    masks = None
    for tag, amount in (("s50", 0.5), ("s55", 0.55)):
        tensors = [sd[n] if masks is None else sd[n].masked_fill(~masks[i], 0) for i, n in enumerate(names)]
        masks, k, thr, ties = O.global_l1_masks(tensors, amount, masks)
        counts = [int((~m).sum()) for m in masks]
        assert counts == g[tag + "_counts"].tolist()
        if ties <= 1:  # tie order of topk is implementation defined
            hashes = [sha16(np.packbits(m.numpy())) for m in masks]
            assert hashes == [str(x) for x in g[tag + "_hashes"]]
    sdm = dict(sd)
    for i, n in enumerate(names):
        sdm[n] = sd[n].masked_fill(~masks[i], 0)
    feat, label, pad = O.synth_batch(2, 200, 80, [200, 150], seed=6)
    with torch.no_grad():
        out = O.model_forward(sdm, cfg, feat, pad, no_pred=True)
    np.testing.assert_allclose(sub(out["hidden"], 10, 32), g["hidden_s55"], atol=2e-4, rtol=1e-4)


def test_kd_loss_terms(golden):
    g = golden("distill")
    tcfg, scfg = base_cfg(20, 12), base_cfg(20, 2)
    for c in (tcfg, scfg):
        c.update(skip_masked=False, skip_nomask=False)
    tsd, ssd = O.synth_state_dict(tcfg, seed=7), O.synth_state_dict(scfg, seed=21)
    feat, label, pad = O.synth_batch(4, 750, 80, LENS)
    for tag, (alpha, T) in {"masked": (0.5, 2.0), "nomasked": (1.0, 1.0)}.items():
        mask = None
        if tag == "masked":
            np.random.seed(1337)
            mask = torch.from_numpy(O.span_mask(4, 750, LENS, 0.7, 5))
        with torch.no_grad():
            t = O.model_forward(tsd, tcfg, feat, pad, label, mask_indices=mask)
            s = O.model_forward(ssd, scfg, feat, pad, label, mask_indices=mask)
        key = "m" if tag == "masked" else "u"
        terms = O.kd_loss(s["logit_" + key], s["label_" + key], t["logit_" + key], T=T, alpha=alpha)
        np.testing.assert_allclose([float(x) for x in terms], g[tag + "_terms"], rtol=2e-5, atol=2e-6)
        assert abs(float(terms[0]) - g[tag + "_loss"][0]) < 2e-5


def test_l1_cosine_matches_torch_ops():
    """this criterion is not in the reference; pin against the PyTorch-op form."""
    g = torch.Generator().manual_seed(3)
    pred, tgt = torch.randn(2, 50, 768, generator=g), torch.randn(2, 50, 768, generator=g)
    tot, l1, cos = O.l1_cosine_loss(pred, tgt, 1.0)
    ref = torch.nn.functional.l1_loss(pred, tgt) + (
        -torch.nn.functional.logsigmoid(torch.nn.functional.cosine_similarity(pred, tgt, dim=-1))).mean()
    assert abs(float(tot) - float(ref)) < 1e-6
