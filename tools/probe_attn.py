"""GPU diagnostic for the fused attention kernels (prints error statistics)."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_ssl_compression_b200 import kernels as K

dev = "cuda"
torch.manual_seed(0)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-9)).item()


def ref_attn(qkv, lens, B, T, H, causal):
    E = 64 * H
    x = qkv.float().view(B, T, 3, H, 64)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    s = (q / 8.0) @ k.transpose(-1, -2)
    ar = torch.arange(T, device=dev)
    kmask = ar[None, :] >= lens[:, None]
    s = s.masked_fill(kmask[:, None, None, :], float("-inf"))
    if causal:
        s = s.masked_fill(torch.ones(T, T, dtype=torch.bool, device=dev).triu(1), float("-inf"))
    p = torch.softmax(s, -1)
    o = (p @ v).transpose(1, 2).reshape(B * T, E)
    return o, p


def run(B, T, H, lens, causal=False, bwd=True):
    E = 64 * H
    qkv = (torch.randn(B * T, 3 * E, device=dev) * 1.0).to(torch.bfloat16)
    lens_t = torch.tensor(lens, device=dev, dtype=torch.int32)
    out, lse, keep = K.attn_fwd(qkv, lens_t, B, T, H, causal=causal)
    torch.cuda.synchronize()
    qr = qkv.float().requires_grad_(True)
    o_ref, _ = ref_attn(qr, lens_t, B, T, H, causal)
    print(f"fwd B={B} T={T} H={H} lens={lens} causal={causal}: rel={rel(out, o_ref):.3e}", flush=True)
    if bwd:
        dout = torch.randn(B * T, E, device=dev).to(torch.bfloat16)
        dqkv = K.attn_bwd(qkv, lens_t, out, dout, lse, keep, B, T, H, causal=causal)
        torch.cuda.synchronize()
        o_ref.backward(dout.float())
        g = qr.grad
        for nm, sl in (("dq", slice(0, E)), ("dk", slice(E, 2 * E)), ("dv", slice(2 * E, 3 * E))):
            print(f"   {nm}: rel={rel(dqkv[:, sl], g[:, sl]):.3e}", flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "fwd"):
    run(1, 128, 1, [128], bwd=False)
    run(2, 256, 2, [256, 200], bwd=False)
    run(2, 300, 3, [300, 77], bwd=False)
if which in ("all", "bwd"):
    run(1, 128, 1, [128])
    run(2, 300, 3, [300, 77])
    run(4, 750, 12, [750, 712, 655, 601])
    run(2, 300, 2, [300, 211], causal=True)
if which in ("all", "drop"):
    B, T, H = 2, 256, 2
    E = 64 * H
    qkv = torch.randn(B * T, 3 * E, device=dev).to(torch.bfloat16)
    lens_t = torch.tensor([256, 256], device=dev, dtype=torch.int32)
    o0, _, _ = K.attn_fwd(qkv, lens_t, B, T, H)
    o1, lse, keep = K.attn_fwd(qkv, lens_t, B, T, H, p_drop=0.1, seed=123, site=7)
    o2, _, _ = K.attn_fwd(qkv, lens_t, B, T, H, p_drop=0.1, seed=123, site=7)
    print("dropout deterministic:", torch.equal(o1, o2), " mean-ratio:", (o1.float().mean() / o0.float().mean()).item(),
          " rel(o1,o0):", rel(o1, o0))
if which in ("all", "perf"):
    for (B, T, H) in [(32, 750, 12), (4, 750, 12), (16, 1500, 12)]:
        E = 64 * H
        qkv = torch.randn(B * T, 3 * E, device=dev).to(torch.bfloat16)
        lens_t = torch.full((B,), T, device=dev, dtype=torch.int32)
        dout = torch.randn(B * T, E, device=dev).to(torch.bfloat16)
        out, lse, keep = K.attn_fwd(qkv, lens_t, B, T, H, p_drop=0.1, seed=1, site=1)
        for nm, fn, mult in (("fwd", lambda: K.attn_fwd(qkv, lens_t, B, T, H, p_drop=0.1, seed=1, site=1), 4),
                             ("fwd_nodrop", lambda: K.attn_fwd(qkv, lens_t, B, T, H), 4),
                             ("bwd", lambda: K.attn_bwd(qkv, lens_t, out, dout, lse, keep, B, T, H, p_drop=0.1, seed=1, site=1), 10)):
            for _ in range(3): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            fl = mult * B * H * T * T * 64
            print(f"perf {nm:10s} B={B} T={T} H={H}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
