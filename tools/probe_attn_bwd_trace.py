"""Debug: clock64 timeline of CTA 0 of the attention backward (library built with -DMH_BWD_TRACE=1)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from speech_ssl_compression_b200 import kernels as K, lib as L
B, T, H = 32, 750, 12
E = 64 * H
qkv = torch.randn(B * T, 3 * E, device="cuda").to(torch.bfloat16)
lens = torch.full((B,), T, device="cuda", dtype=torch.int32)
dout = torch.randn(B * T, E, device="cuda").to(torch.bfloat16)
out, lse, keep = K.attn_fwd(qkv, lens, B, T, H, p_drop=0.1, seed=1, site=1)
for _ in range(3):
    K.attn_bwd(qkv, lens, out, dout, lse, keep, B, T, H, p_drop=0.1, seed=1, site=1)
buf = np.zeros((3, 48, 8), dtype=np.int64)
L.lib().mh_attn_bwd_trace_read(buf.ctypes.data_as(ctypes.c_void_p))
t0 = buf[0, 0, 0]
print("it |  w0: top  wait_S  phaseA  drain  wait_dP  phaseB | w15: top  wait_S  phaseA  drain wait_dP phaseB | mma: P_seen(lag) issue1 wait_dS(lag after w-last) issue2")
for it in range(30):
    a, b, m = buf[0, it], buf[1, it], buf[2, it]
    if a[0] == 0: break
    f = lambda r: f"{r[0]-t0:7d} {r[1]-r[0]:6d} {r[2]-r[1]:6d} {r[3]-r[2]:6d} {r[4]-r[3]:6d} {r[5]-r[4]:6d}"
    last_p = max(a[2], b[2]); last_ds = max(a[5], b[5])
    print(f"{it:2d} | {f(a)} | {f(b)} | {m[0]-t0:7d} ({m[0]-last_p:5d}) {m[1]-m[0]:5d} {m[2]-t0:7d} ({m[2]-last_ds:5d}) {m[3]-m[2]:5d}")
