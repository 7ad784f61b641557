"""Per-iteration timeline of the attention backward (CTA 0: compute thread 0 and the MMA thread), clock64 stamps."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_ssl_compression_b200 import kernels as K, lib

B, T, H = 32, 750, 12
E = 64 * H
qkv = torch.randn(B * T, 3 * E, device="cuda").to(torch.bfloat16)
lens = torch.full((B,), T, device="cuda", dtype=torch.int32)
dout = torch.randn(B * T, E, device="cuda").to(torch.bfloat16)
out, lse, keep = K.attn_fwd(qkv, lens, B, T, H, p_drop=0.1, seed=1, site=1)
os.environ["MH_ATTN_DBG"] = str(256 + (int(sys.argv[1]) if len(sys.argv) > 1 else 0))
for _ in range(2):
    K.attn_bwd(qkv, lens, out, dout, lse, keep, B, T, H, p_drop=0.1, seed=1, site=1)
torch.cuda.synchronize()
L = lib.lib()
buf = (ctypes.c_longlong * 4096)()
assert L.mh_attn_debug_trace(buf, 4096) == 0
names = {0: "c:wait s_full", 1: "c:got s_full", 2: "c:arrived p_ready", 3: "c:wait dp_full", 4: "c:got dp_full", 5: "c:arrived ds_ready",
         6: "c:wait dq_full", 7: "c:got dq_full", 8: "c:drain staged", 9: "c:-", 10: "m:wait p_ready", 11: "m:got p_ready",
         12: "m:wait ds_ready", 13: "m:got ds_ready", 14: "m:issued dQ dK dP", 15: "c:A first tmem ld", 16: "c:A math+sts done", 17: "c:A fences done", 18: "c:B first tmem ld", 19: "c:B math+sts done", 20: "c:B fences done", 21: "c:wait fin_full", 22: "c:got fin_full", 23: "c:arrived acc_free", 24: "c:dK dV stored", 25: "c:next item stats issued"}
ev = []
for base in (0, 2048):
    for i in range(2000):
        v = buf[base + i]
        if v == 0: break
        ev.append((v & 0xffffffffffff, v >> 48))
ev.sort()
lo, hi = int(sys.argv[2]) if len(sys.argv) > 2 else 300, int(sys.argv[3]) if len(sys.argv) > 3 else 420
t0 = ev[lo][0]
prev = t0
for t, s in ev[lo:hi]:
    print(f"{t - t0:8d} (+{t - prev:5d})  {'    ' if 10 <= s <= 14 else ''}{names[s]}")
    prev = t
