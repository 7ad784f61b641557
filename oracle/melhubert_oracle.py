"""CPU oracle: a plain restatement of the reference's MelHuBERT hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.  The
product package (``speech_ssl_compression_b200``) never does; it has no CPU fallback.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here against
fixtures under ``tests/golden/`` that were produced by *executing the unmodified reference*
(``oracle/gen_golden.py``, run in the build container where /root/reference exists).

Everything is functional: a model is just a ``state_dict`` (name -> fp32 tensor, the
reference's parameter names) plus a small config dict.  Float math is fp32 torch-CPU,
integer/bool selection math is numpy/python.  Each function cites the reference lines it
restates (paths relative to the reference root).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

HEAD_DIM = 64


# --------------------------------------------------------------------------------------
# span masks  (fairseq_code/data_utils.py:20-153 as called from model.py:57-84)
# --------------------------------------------------------------------------------------
def span_mask(batch, frames, valid_lens, mask_prob, mask_length, min_masks=2, rng=np.random):
    """HuBERT span mask, the ``static`` / overlapping / ``require_same_masks=False`` branch.

    Consumes the legacy global NumPy stream exactly like the reference: one ``rand()`` up
    front (data_utils.py:57-61), then per row one ``rand()`` (:69-73) and one
    ``choice(sz - min_len, num_mask, replace=False)`` (:129).
    Returns a bool array (batch, frames).
    """
    out = np.zeros((batch, frames), dtype=bool)
    rng.rand()  # drawn (and discarded) even though a padding mask is given
    for b in range(batch):
        sz = int(valid_lens[b])
        n_span = max(min_masks, int(mask_prob * sz / float(mask_length) + rng.rand()))
        min_len = mask_length
        if sz - min_len <= n_span:  # data_utils.py:125-127
            min_len = sz - n_span - 1
        starts = rng.choice(sz - min_len, n_span, replace=False)
        idx = (starts[:, None] + np.arange(mask_length)[None, :]).reshape(-1)
        idx = np.unique(idx[idx < sz])  # :139
        out[b, idx] = True
    return out


def consume_layerdrop_draws(n_layers, rng=np.random):
    """module.py:243 draws ``np.random.random()`` once per layer per forward, even with
    ``encoder_layerdrop == 0`` -- needed for a bit-exact multi-step mask replay."""
    for _ in range(n_layers):
        rng.random()


# --------------------------------------------------------------------------------------
# parameters with optional unstructured masks  (pytorch_code/prune.py:64-85)
# --------------------------------------------------------------------------------------
def eff(sd, name):
    """Effective tensor for ``name``: ``name_orig.masked_fill(~name_mask, 0)`` if pruned."""
    if name in sd:
        return sd[name]
    orig, mask = sd[name + "_orig"], sd[name + "_mask"]
    return orig.masked_fill(~mask.bool(), 0)


def pos_conv_weight(weight_g, weight_v):
    """``nn.utils.weight_norm(conv, dim=2)`` (module.py:187): w = g * v / ||v||, the norm
    taken over dims (0, 1) separately for each of the 128 taps."""
    norm = weight_v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
    return weight_g * weight_v / norm


def gelu_erf(x):
    """fairseq_code/gelu.py:34-35 -- erf GELU evaluated in fp32."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = (x - mu).pow(2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


# --------------------------------------------------------------------------------------
# attention  (pytorch_code/forward_multihead_attention.py:39-69, 113-243)
# --------------------------------------------------------------------------------------
def attention(x, sd, prefix, key_pad, causal=False):
    """x: (B, T, C) fp32.  key_pad: (B, T) bool, True = padded key.  The number of heads is
    read off the (possibly head-pruned) q_proj rows: heads = rows / 64
    (forward_multihead_attention.py:161-166)."""
    B, T, _ = x.shape
    wq, wk, wv = (eff(sd, f"{prefix}.{n}_proj.weight") for n in "qkv")
    bq, bk, bv = (eff(sd, f"{prefix}.{n}_proj.bias") for n in "qkv")
    heads = wq.shape[0] // HEAD_DIM
    q = (x @ wq.t() + bq).view(B, T, heads, HEAD_DIM).transpose(1, 2)
    k = (x @ wk.t() + bk).view(B, T, heads, HEAD_DIM).transpose(1, 2)
    v = (x @ wv.t() + bv).view(B, T, heads, HEAD_DIM).transpose(1, 2)
    s = (q / math.sqrt(HEAD_DIM)) @ k.transpose(-1, -2)  # (B, h, T, T)
    neg = torch.zeros(B, 1, 1, T, device=x.device)  # fp32 additive mask (forward_multihead_attention.py:221)
    neg.masked_fill_(key_pad.view(B, 1, 1, T), float("-inf"))
    s = s + neg
    if causal:
        tri = torch.ones(T, T, dtype=torch.bool, device=x.device).triu(1)
        s = s.masked_fill(tri, float("-inf"))
    p = torch.softmax(s, dim=-1)
    ctx = (p @ v).transpose(1, 2).reshape(B, T, heads * HEAD_DIM)
    wo, bo = eff(sd, f"{prefix}.out_proj.weight"), eff(sd, f"{prefix}.out_proj.bias")
    return ctx @ wo.t() + bo


# --------------------------------------------------------------------------------------
# encoder  (module.py:82-133 post-LN / pre-LN block, :225-257 extract_features)
# --------------------------------------------------------------------------------------
def encoder_layer(x, sd, prefix, key_pad, layer_norm_first=False, causal=False):
    ln1 = (sd[f"{prefix}.self_attn_layer_norm.weight"], sd[f"{prefix}.self_attn_layer_norm.bias"])
    ln2 = (sd[f"{prefix}.final_layer_norm.weight"], sd[f"{prefix}.final_layer_norm.bias"])
    w1, b1 = eff(sd, f"{prefix}.fc1.weight"), eff(sd, f"{prefix}.fc1.bias")
    w2, b2 = eff(sd, f"{prefix}.fc2.weight"), eff(sd, f"{prefix}.fc2.bias")
    if layer_norm_first:
        x = x + attention(layer_norm(x, *ln1), sd, f"{prefix}.self_attn", key_pad, causal)
        h = gelu_erf(layer_norm(x, *ln2) @ w1.t() + b1)
        return x + (h @ w2.t() + b2)
    x = layer_norm(x + attention(x, sd, f"{prefix}.self_attn", key_pad, causal), *ln1)
    h = gelu_erf(x @ w1.t() + b1)
    return layer_norm(x + (h @ w2.t() + b2), *ln2)


def encoder(x, sd, key_pad, n_layers, layer_norm_first=False, causal=False, groups=16):
    """x (B,T,C) -> (hidden, [per-layer hiddens]).  Dropout is off (eval / p = 0)."""
    x = x.clone()
    x[key_pad] = 0  # module.py:226-227 (in place on pre_feat in the reference)
    pre = x
    w = pos_conv_weight(sd["encoder.pos_conv.0.weight_g"], sd["encoder.pos_conv.0.weight_v"])
    k = w.shape[-1]
    pc = F.conv1d(x.transpose(1, 2), w, sd["encoder.pos_conv.0.bias"], padding=k // 2, groups=groups)
    if k % 2 == 0:
        pc = pc[:, :, :-1]  # SamePad (same_pad.py:17-28)
    x = x + gelu_erf(pc).transpose(1, 2)
    if not layer_norm_first:
        x = layer_norm(x, sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"])
    hiddens = []
    for i in range(n_layers):
        x = encoder_layer(x, sd, f"encoder.layers.{i}", key_pad, layer_norm_first, causal)
        hiddens.append(x)
    if layer_norm_first:
        x = layer_norm(x, sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"])
    return x, hiddens, pre


# --------------------------------------------------------------------------------------
# model forward  (model.py:86-163)
# --------------------------------------------------------------------------------------
def model_forward(sd, cfg, feat, pad_mask, label=None, mask_indices=None, no_pred=False):
    """cfg keys used: encoder_layers, layer_norm_first, attention_type, skip_masked,
    skip_nomask.  ``mask_indices`` (B,T) bool or None (= no masking; model.py:103-104).
    Masking before projection with mask_emb = 0 (the shipped configuration).
    Returns a dict with the reference tuple's members."""
    feat = feat.clone()
    valid = pad_mask.bool()
    if mask_indices is None:
        mask_indices = torch.zeros_like(valid)
    else:
        feat[mask_indices] = 0  # model.py:80
    pre = feat @ sd["pre_extract_proj.weight"].t() + sd["pre_extract_proj.bias"]
    hidden, hiddens, pre_feat = encoder(
        pre, sd, ~valid, int(cfg["encoder_layers"]), bool(cfg.get("layer_norm_first", False)),
        cfg.get("attention_type", "original") == "causal")
    out = dict(hidden=hidden, layer_hiddens=hiddens, pre_feat=pre_feat, mask_indices=mask_indices)
    if no_pred:
        return out
    wf, bf = sd["final_proj.weight"], sd["final_proj.bias"]
    if not cfg.get("skip_masked", False):
        sel = valid & mask_indices  # row-major (b, t) order, model.py:147-150
        out["logit_m"] = hidden[sel] @ wf.t() + bf
        out["label_m"] = label[sel]
    if not cfg.get("skip_nomask", True):
        sel = valid & ~mask_indices
        out["logit_u"] = hidden[sel] @ wf.t() + bf
        out["label_u"] = label[sel]
    return out


# --------------------------------------------------------------------------------------
# criteria
# --------------------------------------------------------------------------------------
def ce_mean(logits, labels, ignore_index=-100):
    """upstream/melhubert/pretrain_expert.py:25,116-117 -- CrossEntropyLoss(mean, ignore -100)."""
    keep = labels != ignore_index
    lse = torch.logsumexp(logits, dim=-1)
    picked = logits.gather(1, labels.clamp(min=0).view(-1, 1)).squeeze(1)
    return ((lse - picked) * keep).sum() / keep.sum()


def kd_loss(student, labels, teacher, T=1.0, alpha=0.5):
    """distillation/pretrain_expert.py:83-92 -- (1-a)*CE + a*KL_batchmean(log_softmax(s/T) || softmax(t/T)).
    Returns (total, hard, soft, teacher_ce)."""
    hard = ce_mean(student, labels)
    t_ce = ce_mean(teacher, labels)
    ls = torch.log_softmax(student / T, dim=1)
    lt = torch.log_softmax(teacher / T, dim=1)
    soft = (lt.exp() * (lt - ls)).sum() / student.shape[0]
    return hard * (1.0 - alpha) + soft * alpha, hard, soft, t_ce


def l1_cosine_loss(pred, target, cos_weight=1.0):
    """north_star's per-layer L1 + cosine criterion (DistilHuBERT form).  NOT in the
    reference tree -- parity for this one function is pinned against the
    PyTorch-op restatement ``F.l1_loss + w * (-logsigmoid(cosine_similarity)).mean()``
    (see tests), not against reference output.
    pred/target: (L, N, C).  Returns (total, l1, cos)."""
    l1 = (pred - target).abs().mean()
    num = (pred * target).sum(-1)
    den = pred.norm(dim=-1).clamp(min=1e-8) * target.norm(dim=-1).clamp(min=1e-8)
    cos = num / den
    cos_term = F.softplus(-cos).mean()  # = -logsigmoid(cos)
    return l1 + cos_weight * cos_term, l1, cos_term


# --------------------------------------------------------------------------------------
# pruning objects (bit-exact selections)
# --------------------------------------------------------------------------------------
PRUNE_ORDER = ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.out_proj", "fc1", "fc2")


def prunable_names(n_layers, bias=True):
    """weight_pruning/wp_utils.py:13-48: per layer the six weights, then the six biases."""
    names = []
    for i in range(n_layers):
        names += [f"encoder.layers.{i}.{m}.weight" for m in PRUNE_ORDER]
        if bias:
            names += [f"encoder.layers.{i}.{m}.bias" for m in PRUNE_ORDER]
    return names


def global_l1_masks(tensors, amount, old_masks=None):
    """pytorch_code/prune.py:1049-1171 + :553-573.  ``tensors``: list of effective (already
    masked) tensors in ``prunable_names`` order.  k = int(round(amount * N)) smallest |w|
    over the concatenation are masked (``topk(largest=False)``), AND-ed with old masks.
    Returns (list of bool masks, k, threshold, n_ties_at_threshold)."""
    flat = torch.cat([t.reshape(-1) for t in tensors]).abs()
    n = flat.numel()
    k = int(round(amount * n))
    mask = torch.ones(n, dtype=torch.bool) if old_masks is None else torch.cat(
        [m.reshape(-1) for m in old_masks]).clone()
    thr, ties = None, 0
    if k:
        top = flat.topk(k, largest=False)
        mask[top.indices] = False
        thr = float(top.values.max())
        ties = int((flat == thr).sum())
    out, p = [], 0
    for t in tensors:
        out.append(mask[p:p + t.numel()].view_as(t))
        p += t.numel()
    return out, k, thr, ties


def head_scores(sd, layer, head_dim=HEAD_DIM):
    """head_pruning/hp_utils.py:188-232: score(h) = sum|Wk[h]| + sum|bk[h]| + (q) + (v), each
    term an fp32 torch.sum converted to a Python float, added as doubles in k, q, v order."""
    pre = f"encoder.layers.{layer}.self_attn"
    heads = sd[f"{pre}.q_proj.weight"].shape[0] // head_dim
    out = []
    for h in range(heads):
        sl = slice(h * head_dim, (h + 1) * head_dim)
        parts = []
        for n in ("k", "q", "v"):
            w, b = sd[f"{pre}.{n}_proj.weight"], sd[f"{pre}.{n}_proj.bias"]
            parts.append(torch.sum(torch.abs(w[sl])).tolist() + torch.sum(torch.abs(b[sl])).tolist())
        out.append(parts[0] + parts[1] + parts[2])
    return out


def select_heads(scores_per_layer, target="by_layer", n_to_prune=None):
    """hp_utils.py:53-106.  ``scores_per_layer``: list (layer) of list (head) of float.
    Returns {layer: [head, ...]} in the reference's insertion order."""
    flat = [((l, h), s) for l, row in enumerate(scores_per_layer) for h, s in enumerate(row)]
    ranked = [lh for lh, _ in sorted(flat, key=lambda x: x[1])]  # stable
    n_layers = len(scores_per_layer)
    if n_to_prune is None:
        n_to_prune = n_layers
    if target == "by_whole":
        protect = {l: 1 for l in range(n_layers)}
        kept = []
        for l, h in reversed(ranked):
            if l in protect:
                if protect[l] > 0:
                    protect[l] -= 1
                    continue
                protect.pop(l)
            kept.insert(0, (l, h))
        chosen = kept[:n_to_prune]
    else:
        want = set(range(n_to_prune))
        chosen = []
        for l, h in ranked:
            if not want:
                break
            if l in want:
                chosen.append((l, h))
                want.remove(l)
    group = {}
    for l, h in chosen:
        group[l] = group.get(l, []) + [h]
    return group


def slice_heads(sd, layer, heads, head_dim=HEAD_DIM):
    """hp_utils.py:108-186: drop the given heads' rows of q/k/v (+bias) and columns of out_proj."""
    pre = f"encoder.layers.{layer}.self_attn"
    n = sd[f"{pre}.q_proj.weight"].shape[0] // head_dim
    keep = torch.cat([torch.arange(h * head_dim, (h + 1) * head_dim) for h in range(n) if h not in heads])
    for p in "qkv":
        sd[f"{pre}.{p}_proj.weight"] = sd[f"{pre}.{p}_proj.weight"][keep].clone()
        sd[f"{pre}.{p}_proj.bias"] = sd[f"{pre}.{p}_proj.bias"][keep].clone()
    sd[f"{pre}.out_proj.weight"] = sd[f"{pre}.out_proj.weight"][:, keep].clone()


def row_scores(sd, layer):
    """row_pruning/rp_utils.py:84-112: score(i) = (sum|W1[i,:]| + |b1[i]|) + sum|W2[:,i]|."""
    w1, b1 = sd[f"encoder.layers.{layer}.fc1.weight"], sd[f"encoder.layers.{layer}.fc1.bias"]
    w2 = sd[f"encoder.layers.{layer}.fc2.weight"]
    out = []
    for i in range(w1.shape[0]):
        a = torch.sum(torch.abs(w1[i, :])).tolist() + torch.abs(b1[i]).tolist()
        out.append(a + torch.sum(torch.abs(w2[:, i])).tolist())
    return out


def select_rows(scores, n_to_prune):
    """rp_utils.py:40-48: the ``n_to_prune`` lowest-score rows (stable sort)."""
    ranked = sorted(range(len(scores)), key=lambda i: scores[i])
    return ranked[:n_to_prune]


def slice_rows(sd, layer, rows):
    """rp_utils.py:50-82."""
    pre = f"encoder.layers.{layer}"
    drop = set(rows)
    keep = torch.tensor([i for i in range(sd[f"{pre}.fc1.weight"].shape[0]) if i not in drop])
    sd[f"{pre}.fc1.weight"] = sd[f"{pre}.fc1.weight"][keep].clone()
    sd[f"{pre}.fc1.bias"] = sd[f"{pre}.fc1.bias"][keep].clone()
    sd[f"{pre}.fc2.weight"] = sd[f"{pre}.fc2.weight"][:, keep].clone()


# --------------------------------------------------------------------------------------
# deterministic synthetic weights / batches shared by the golden generator and the tests
# --------------------------------------------------------------------------------------
def synth_state_dict(cfg, seed=7, heads_per_layer=None, ffn_per_layer=None):
    """Weights that do not depend on nn.Module construction order: every tensor is drawn
    from its own ``torch.Generator`` seeded by (seed, crc32(name)).  Statistics mimic the
    reference init (N(0, 0.02) linears, LN = (1, 0)) but biases and LN affine terms are
    made non-trivial so that parity tests exercise them."""
    import zlib

    d = int(cfg.get("encoder_embed_dim", 768))
    f = int(cfg.get("encoder_ffn_embed_dim", 3072))
    n_layers = int(cfg["encoder_layers"])
    d_in = int(cfg.get("feat_emb_dim", 40))
    k = int(cfg.get("num_cluster", 512))
    taps, groups = int(cfg.get("conv_pos", 128)), int(cfg.get("conv_pos_groups", 16))
    h0 = int(cfg.get("encoder_attention_heads", 12))

    def draw(name, shape, std, mean=0.0):
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
        return torch.randn(shape, generator=g) * std + mean

    sd = {}
    sd["pre_extract_proj.weight"] = draw("pre_extract_proj.weight", (d, d_in), 0.05)
    sd["pre_extract_proj.bias"] = draw("pre_extract_proj.bias", (d,), 0.02)
    sd["encoder.pos_conv.0.bias"] = draw("encoder.pos_conv.0.bias", (d,), 0.02)
    sd["encoder.pos_conv.0.weight_g"] = draw("encoder.pos_conv.0.weight_g", (1, 1, taps), 0.05, 0.6)
    sd["encoder.pos_conv.0.weight_v"] = draw("encoder.pos_conv.0.weight_v", (d, d // groups, taps), 0.02)
    for i in range(n_layers):
        e = HEAD_DIM * (heads_per_layer[i] if heads_per_layer else h0)
        fl = ffn_per_layer[i] if ffn_per_layer else f
        p = f"encoder.layers.{i}"
        for n in "qkv":
            sd[f"{p}.self_attn.{n}_proj.weight"] = draw(f"{p}.{n}.w", (e, d), 0.03)
            sd[f"{p}.self_attn.{n}_proj.bias"] = draw(f"{p}.{n}.b", (e,), 0.02)
        sd[f"{p}.self_attn.out_proj.weight"] = draw(f"{p}.o.w", (d, e), 0.03)
        sd[f"{p}.self_attn.out_proj.bias"] = draw(f"{p}.o.b", (d,), 0.02)
        sd[f"{p}.self_attn_layer_norm.weight"] = draw(f"{p}.ln1.w", (d,), 0.05, 1.0)
        sd[f"{p}.self_attn_layer_norm.bias"] = draw(f"{p}.ln1.b", (d,), 0.02)
        sd[f"{p}.fc1.weight"] = draw(f"{p}.fc1.w", (fl, d), 0.03)
        sd[f"{p}.fc1.bias"] = draw(f"{p}.fc1.b", (fl,), 0.02)
        sd[f"{p}.fc2.weight"] = draw(f"{p}.fc2.w", (d, fl), 0.03)
        sd[f"{p}.fc2.bias"] = draw(f"{p}.fc2.b", (d,), 0.02)
        sd[f"{p}.final_layer_norm.weight"] = draw(f"{p}.ln2.w", (d,), 0.05, 1.0)
        sd[f"{p}.final_layer_norm.bias"] = draw(f"{p}.ln2.b", (d,), 0.02)
    sd["encoder.layer_norm.weight"] = draw("encoder.layer_norm.weight", (d,), 0.05, 1.0)
    sd["encoder.layer_norm.bias"] = draw("encoder.layer_norm.bias", (d,), 0.02)
    sd["final_proj.weight"] = draw("final_proj.weight", (k, d), 0.03)
    sd["final_proj.bias"] = draw("final_proj.bias", (k,), 0.02)
    return sd


def synth_batch(batch, frames, d_in, lens, seed=2024, n_cluster=512):
    """SURVEY appendix C recipe: feat ~ N(0,1), labels U{0..K-1}; suffix padding with
    feat = 0, label = -100, pad_mask = 0."""
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(batch, frames, d_in, generator=g)
    label = torch.randint(0, n_cluster, (batch, frames), generator=g)
    pad = torch.ones(batch, frames)
    for i, l in enumerate(lens):
        pad[i, l:] = 0
        label[i, l:] = -100
        feat[i, l:] = 0
    return feat, label, pad
