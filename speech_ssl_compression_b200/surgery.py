"""Structural edits of the module tree shared by the experts, the pruning tools and
``extract_feature.py``: shrinking attention heads / FFN rows (reference
``head_pruning/hp_utils.py:108-186``, ``row_pruning/rp_utils.py:50-82``,
``upstream/melhubert/pretrain_expert.py:45-65``, ``extract_feature.py:116-138``)."""
import torch
import torch.nn as nn


def summarize_pruned_heads(record):
    """``Pruned_heads`` ckpt entry (list of {layer: [head, ...]} per prune step) -> {layer: count}."""
    out = {}
    for step in record:
        for layer, heads in step.items():
            out[layer] = out.get(layer, 0) + len(heads)
    return out


def rebuild_attention_for_heads(mha, n_removed):
    """Re-create q/k/v/out projections for an attention module that lost ``n_removed`` heads
    (shapes only; weights come from the checkpoint afterwards)."""
    mha.num_heads -= n_removed
    full = mha.embed_dim
    small = mha.head_dim * mha.num_heads
    mha.embed_dim = small
    dev = mha.out_proj.weight.device
    mha.k_proj = nn.Linear(full, small, bias=True).to(dev)
    mha.v_proj = nn.Linear(full, small, bias=True).to(dev)
    mha.q_proj = nn.Linear(full, small, bias=True).to(dev)
    mha.out_proj = nn.Linear(small, full, bias=True).to(dev)
    mha.skip_embed_dim_check = True
    mha.reset_parameters()
    mha.__dict__.pop("_mh_operands", None)


def apply_pruned_heads_record(model, record):
    for idx, layer in enumerate(model.encoder.layers):
        n = summarize_pruned_heads(record).get(idx, 0)
        if n:
            rebuild_attention_for_heads(layer.self_attn, n)


def _new_param(t):
    return nn.Parameter(t.detach().clone().contiguous(), requires_grad=True)


def drop_heads(mha, heads):
    """Physically remove ``heads`` (indices in the current layout): rows of q/k/v (+ bias),
    columns of out_proj.  ``out_proj.bias`` is untouched."""
    hd = mha.head_dim
    keep = [h for h in range(mha.num_heads) if h not in set(heads)]
    rows = torch.cat([torch.arange(h * hd, (h + 1) * hd) for h in keep]).to(mha.q_proj.weight.device)
    for proj in (mha.q_proj, mha.k_proj, mha.v_proj):
        proj.weight = _new_param(proj.weight[rows])
        proj.bias = _new_param(proj.bias[rows])
    mha.out_proj.weight = _new_param(mha.out_proj.weight[:, rows])
    mha.num_heads = len(keep)
    mha.embed_dim = hd * mha.num_heads
    for proj in (mha.q_proj, mha.k_proj, mha.v_proj):
        proj.out_features = mha.embed_dim
    mha.out_proj.in_features = mha.embed_dim
    mha._set_skip_embed_dim_check()
    mha.__dict__.pop("_mh_operands", None)


def drop_ffn_rows(layer, rows):
    fc1, fc2 = layer.fc1, layer.fc2
    gone = set(int(r) for r in rows)
    keep = torch.tensor([i for i in range(fc1.weight.shape[0]) if i not in gone], device=fc1.weight.device)
    fc1.weight = _new_param(fc1.weight[keep])
    fc1.bias = _new_param(fc1.bias[keep])
    fc2.weight = _new_param(fc2.weight[:, keep])
    fc1.out_features = fc2.in_features = int(keep.numel())
    layer.__dict__.pop("_mh_operands", None)
