"""Tiny driver for ncu: three forward (dropout 0.1) + backward launches of the fused attention at the
bench shape (B=32, T=750, 12 heads)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_ssl_compression_b200 import kernels as K

B, T, H = 32, 750, 12
E = 64 * H
torch.manual_seed(0)
qkv = torch.randn(B * T, 3 * E, device="cuda").to(torch.bfloat16)
lens = torch.full((B,), T, device="cuda", dtype=torch.int32)
dout = torch.randn(B * T, E, device="cuda").to(torch.bfloat16)
for _ in range(3):
    out, lse, keep = K.attn_fwd(qkv, lens, B, T, H, p_drop=0.1, seed=1, site=1)
    K.attn_bwd(qkv, lens, out, dout, lse, keep, B, T, H, p_drop=0.1, seed=1, site=1)
torch.cuda.synchronize()
print("ok")
