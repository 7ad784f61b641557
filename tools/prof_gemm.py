"""Tiny driver for ncu: the fc1 forward GEMM (bias + GELU + dropout epilogue, two outputs) and the plain
QKV GEMM at the bench shape (M = 24000)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_ssl_compression_b200 import kernels as K

M = 24000
torch.manual_seed(0)
x = torch.randn(M, 768, device="cuda").to(torch.bfloat16)
w1 = torch.randn(3072, 768, device="cuda").to(torch.bfloat16)
b1 = torch.zeros(3072, device="cuda")
u = torch.empty(M, 3072, device="cuda", dtype=torch.bfloat16)
pre = torch.empty_like(u)
wq = torch.randn(2304, 768, device="cuda").to(torch.bfloat16)
bq = torch.zeros(2304, device="cuda")
qkv = torch.empty(M, 2304, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    K.gemm(x, w1, u, epilogue=K.EPI_GELU, bias=b1, aux_out=pre, p_drop=0.1, seed=1, site=2)
    K.gemm(x, wq, qkv, bias=bq)
torch.cuda.synchronize()
print("ok")
