"""Pruning re-parametrisation (the subset of the reference's ``pytorch_code`` package that the
MelHuBERT path uses).  The attention math of the reference's
``pytorch_code/forward_multihead_attention.py`` lives in the fused CUDA kernels instead
(``csrc/attn_sm100.cu``, ``csrc/gemm_sm100.cu``)."""
from . import prune  # noqa: F401
