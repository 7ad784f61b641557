"""GPU diagnostic for the tcgen05 GEMM: prints error statistics per variant (does not assert)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_ssl_compression_b200 import kernels as K

torch.manual_seed(0)
dev = "cuda"


def report(name, got, ref):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / ref.norm().clamp_min(1e-9)).item()
    print(f"{name:58s} max_abs={err:.4e} rel_l2={rel:.3e} {'OK' if rel < 1e-2 else 'BAD'}", flush=True)
    return rel


def run(M, N, Kd, a_mn, b_mn, bn=0, epi=K.EPI_BF16):
    a = torch.randn(M, Kd, device=dev).to(torch.bfloat16)
    b = torch.randn(N, Kd, device=dev).to(torch.bfloat16)
    ref = a.float() @ b.float().t()
    a_st = a.t().contiguous() if a_mn else a
    b_st = b.t().contiguous() if b_mn else b
    if epi == K.EPI_F32:
        out = torch.zeros(M, N, device=dev, dtype=torch.float32)
    else:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    K.gemm(a_st, b_st, out, a_mn=a_mn, b_mn=b_mn, epilogue=epi, block_n=bn)
    torch.cuda.synchronize()
    return report(f"M={M} N={N} K={Kd} a_mn={int(a_mn)} b_mn={int(b_mn)} bn={bn} epi={epi}", out, ref)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "kmajor"):
    run(128, 128, 64, False, False, 128)
    run(128, 128, 256, False, False, 128)
    run(256, 256, 768, False, False, 256)
    run(3000, 2304, 768, False, False)
    run(3000, 768, 3072, False, False)
    run(1000, 2112, 80, False, False)
    run(333, 64, 40, False, False)
if which in ("all", "mn"):
    run(128, 128, 64, True, False, 128)
    run(128, 128, 64, False, True, 128)
    run(128, 128, 64, True, True, 128)
    run(256, 256, 256, True, False, 256)
    run(256, 256, 256, False, True, 256)
    run(768, 3072, 3000, True, True, 0, K.EPI_F32)
    run(768, 768, 3000, True, True, 0, K.EPI_F32)
    run(2304, 768, 24000, True, True, 0, K.EPI_F32)
if which in ("all", "perf"):
    for (M, N, Kd) in [(24000, 2304, 768), (24000, 3072, 768), (24000, 768, 3072), (24000, 768, 768), (3000, 3072, 768)]:
        a = torch.randn(M, Kd, device=dev).to(torch.bfloat16)
        b = torch.randn(N, Kd, device=dev).to(torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for fn, nm in ((lambda: K.gemm(a, b, out), "mh"), (lambda: torch.matmul(a, b.t(), out=out), "torch")):
            for _ in range(3): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"perf {nm:6s} M={M} N={N} K={Kd}: {ms*1e3:.1f} us  {2*M*N*Kd/ms/1e9:.1f} TFLOP/s", flush=True)
