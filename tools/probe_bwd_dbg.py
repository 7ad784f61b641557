"""Timing experiments on the attention backward (MH_ATTN_DBG disables parts of the kernel; results are garbage,
only the durations mean anything)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_ssl_compression_b200 import kernels as K

B, T, H = 32, 750, 12
E = 64 * H
qkv = torch.randn(B * T, 3 * E, device="cuda").to(torch.bfloat16)
lens = torch.full((B,), T, device="cuda", dtype=torch.int32)
dout = torch.randn(B * T, E, device="cuda").to(torch.bfloat16)
out, lse, keep = K.attn_fwd(qkv, lens, B, T, H, p_drop=0.1, seed=1, site=1)
for dbg in [int(a) for a in sys.argv[1:]] or [0]:
    os.environ["MH_ATTN_DBG"] = str(dbg)
    fn = lambda: K.attn_bwd(qkv, lens, out, dout, lse, keep, B, T, H, p_drop=0.1, seed=1, site=1)
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"dbg={dbg:2d}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us (whole mh_attn_bwd: delta + main + dq_finish)", flush=True)
