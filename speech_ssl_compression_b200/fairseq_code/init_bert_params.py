"""BERT-style re-initialisation (reference ``fairseq_code/init_bert_params.py:19-50``): every
``nn.Linear`` weight and the q/k/v projections of every attention module are redrawn from
N(0, 0.02) **on the CPU generator** (so random-init weights do not depend on the device),
Linear biases are zeroed."""
import torch.nn as nn


def _redraw(t):
    t.copy_(t.cpu().normal_(mean=0.0, std=0.02).to(t.device))


def init_bert_params(module):
    from .multihead_attention import MultiheadAttention

    if isinstance(module, nn.Linear):
        _redraw(module.weight.data)
        if module.bias is not None:
            module.bias.data.zero_()
    if isinstance(module, nn.Embedding):
        _redraw(module.weight.data)
        if module.padding_idx is not None:
            module.weight.data[module.padding_idx].zero_()
    if isinstance(module, MultiheadAttention):
        for proj in (module.q_proj, module.k_proj, module.v_proj):
            _redraw(proj.weight.data)
