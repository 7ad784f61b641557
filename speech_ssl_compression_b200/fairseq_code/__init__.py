"""The pieces of the reference's ``fairseq_code`` package that the MelHuBERT path touches,
re-hosted on the fused sm_100a kernels."""
from .data_utils import compute_mask_indices  # noqa: F401
from .multihead_attention import MultiheadAttention, FairseqDropout  # noqa: F401
from .init_bert_params import init_bert_params  # noqa: F401
