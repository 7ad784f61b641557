"""GPU diagnostic for the positional-conv kernels vs torch conv1d (prints errors; times them)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from speech_ssl_compression_b200 import kernels as K

dev = "cuda"
torch.manual_seed(0)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def run(B, T, C=768, groups=16):
    cg = C // groups
    v = torch.randn(C, cg, 128, device=dev) * 0.02
    g = torch.rand(1, 1, 128, device=dev) + 0.5
    bias = torch.randn(C, device=dev) * 0.1
    x = torch.randn(B * T, C, device=dev).to(torch.bfloat16)
    w_fwd, w_bwd, norm = K.posconv_weight_prep(v, g)
    y, z = K.posconv_fwd(x, w_fwd, bias, B, T)
    torch.cuda.synchronize()
    vr = v.clone().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    w = torch._weight_norm(vr, gr, 2)
    xr = x.float().view(B, T, C).requires_grad_(True)
    zr = F.conv1d(xr.transpose(1, 2), w, bias, padding=64, groups=groups)[:, :, :-1].transpose(1, 2)
    yr = xr + F.gelu(zr)
    print(f"B={B} T={T}: norm rel {rel(norm, v.pow(2).sum((0,1)).sqrt()):.2e}  z rel {rel(z, zr.reshape(B*T, C)):.3e}  y rel {rel(y, yr.reshape(B*T, C)):.3e}", flush=True)
    dy = torch.randn(B * T, C, device=dev).to(torch.bfloat16)
    yr.backward(dy.float().view(B, T, C))
    dz = K.gelu_bwd_mul(dy, z)
    dx = K.posconv_dgrad(dz, w_bwd, dy, B, T)
    torch.cuda.synchronize()
    print(f"   dx rel {rel(dx, xr.grad.reshape(B*T, C)):.3e}", flush=True)
    # reference dW w.r.t. the effective weight
    w2 = w.detach().clone().requires_grad_(True)
    z2 = F.conv1d(x.float().view(B, T, C).transpose(1, 2), w2, bias, padding=64, groups=groups)[:, :, :-1].transpose(1, 2)
    z2.backward(dz.float().view(B, T, C))
    dw = torch.zeros(C, cg, 128, device=dev)
    K.posconv_wgrad(dz, x, dw, B, T)
    torch.cuda.synchronize()
    print(f"   wgrad rel {rel(dw, w2.grad):.3e}", flush=True)
    dv, dg = torch.zeros_like(v), torch.zeros(128, device=dev)
    K.posconv_weight_bwd(dw, v, g.reshape(-1), norm, dv, dg)
    print(f"   dv rel {rel(dv, vr.grad):.3e}  dg rel {rel(dg, gr.grad.reshape(-1)):.3e}", flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "check"):
    run(1, 128)
    run(2, 300)
    run(4, 750)
    run(2, 1500)
if which in ("all", "perf"):
    B, T, C = 32, 750, 768
    v = torch.randn(C, 48, 128, device=dev) * 0.02
    g = torch.rand(1, 1, 128, device=dev) + 0.5
    bias = torch.zeros(C, device=dev)
    x = torch.randn(B * T, C, device=dev).to(torch.bfloat16)
    dw = torch.zeros(C, 48, 128, device=dev)
    w_fwd, w_bwd, norm = K.posconv_weight_prep(v, g)
    y, z = K.posconv_fwd(x, w_fwd, bias, B, T)
    fl = 2.0 * B * T * C * 48 * 128
    for nm, fn in (("prep", lambda: K.posconv_weight_prep(v, g)), ("fwd", lambda: K.posconv_fwd(x, w_fwd, bias, B, T)),
                   ("dgrad", lambda: K.posconv_dgrad(z, w_bwd, x, B, T)), ("wgrad", lambda: K.posconv_wgrad(z, x, dw, B, T)),
                   ("gelu_bwd", lambda: K.gelu_bwd_mul(x, z))):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"perf {nm:9s}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s (conv flops)", flush=True)
