"""Minimal launcher for ncu captures of the attention kernels: python tools/probe_attn_one.py fwd|bwd [B T H p_drop]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_ssl_compression_b200 import kernels as K

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
B, T, H = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (32, 750, 12)
p = float(sys.argv[5]) if len(sys.argv) > 5 else 0.1
n = int(sys.argv[6]) if len(sys.argv) > 6 else 3
E = 64 * H
torch.manual_seed(0)
qkv = torch.randn(B * T, 3 * E, device="cuda").to(torch.bfloat16)
lens = torch.full((B,), T, device="cuda", dtype=torch.int32)
dout = torch.randn(B * T, E, device="cuda").to(torch.bfloat16)
out, lse, keep = K.attn_fwd(qkv, lens, B, T, H, p_drop=p, seed=1, site=1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(2):
    e0.record()
    for _ in range(n):
        if which == "fwd":
            K.attn_fwd(qkv, lens, B, T, H, p_drop=p, seed=1, site=1)
        else:
            K.attn_bwd(qkv, lens, out, dout, lse, keep, B, T, H, p_drop=p, seed=1, site=1)
    e1.record()
    torch.cuda.synchronize()
print(f"{which} B={B} T={T} H={H} p={p}: {e0.elapsed_time(e1) / n * 1e3:.1f} us")
