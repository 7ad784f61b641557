"""MelHuBERT model configuration (reference ``model_config.py:1-47``): a plain attribute bag
filled from the ``melhubert`` section of the model yaml, with the reference's defaults."""

# field -> (type, default); order and defaults follow the reference's MelHuBERTConfig
_FIELDS = (
    ("feat_emb_dim", int, 40),
    ("pos_emb_type", str, "conv"), ("pos_conv_depth", int, 1), ("conv_pos", int, 128), ("conv_pos_groups", int, 16),
    ("encoder_layers", int, 1), ("encoder_embed_dim", int, 768), ("encoder_ffn_embed_dim", int, 3072),
    ("encoder_attention_heads", int, 12), ("activation_fn", str, "gelu"), ("layer_norm_first", bool, False),
    ("attention_type", str, "original"),
    ("num_cluster", int, 512), ("final_dim", int, 40),
    ("pred_masked_weight", float, 1.0), ("pred_nomask_weight", float, 0.0),
    ("mask_prob", float, 0.8), ("mask_length", int, 10), ("mask_selection", str, "static"), ("mask_other", float, 0.0),
    ("no_mask_overlap", bool, False), ("mask_min_space", int, 1),
    ("skip_masked", bool, False), ("skip_nomask", bool, True),
    ("learnable_mask_emb", bool, False), ("mask_before_proj", bool, True),
    ("dropout", float, 0.1), ("attention_dropout", float, 0.1), ("activation_dropout", float, 0.1),
    ("encoder_layerdrop", float, 0.0),
)


class MelHuBERTConfig:
    def __init__(self, config: dict):
        for name, typ, default in _FIELDS:
            setattr(self, name, typ(config.get(name, default)))

    def to_dict(self):
        return {name: getattr(self, name) for name, _, _ in _FIELDS}

    def __repr__(self):
        return f"MelHuBERTConfig({self.to_dict()})"
