"""Debug: print the clock64 timeline of traced CTAs of the forward kernel (library built with -DMH_F2_TRACE=1)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from speech_ssl_compression_b200 import kernels as K, lib as L
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
B, T, H = 32, 750, 12
E = 64 * H
qkv = torch.randn(B * T, 3 * E, device="cuda").to(torch.bfloat16)
lens = torch.full((B,), T, device="cuda", dtype=torch.int32)
for _ in range(3):
    K.attn_fwd(qkv, lens, B, T, H, p_drop=p, seed=1, site=1)
buf = np.zeros((8, 64, 8), dtype=np.int64)
L.lib().mh_attn_trace_read(buf.ctypes.data_as(ctypes.c_void_p))
for c in range(8):
    t0 = buf[c, 63, 0]
    print(f"CTA {c}: start 0, softmax-end {buf[c,63,1]-t0}, exit {buf[c,63,2]-t0}")
    prev = t0
    for j in range(26):
        r = buf[c, j]
        if r[0] == 0: break
        print(f"  blk {j:2d}: top {r[0]-t0:7d} | wait_s {r[1]-r[0]:5d} exp {r[2]-r[1]:5d} drop {r[3]-r[2]:5d} st+arrive {r[4]-r[3]:5d} | total {r[4]-r[0]:5d} || mma: p_seen {r[5]-t0:7d} (lag {r[5]-r[4]:5d}) issue {r[6]-r[5]:5d}")
