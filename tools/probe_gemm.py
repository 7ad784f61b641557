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

if which in ("all", "epi"):
    M = 24000
    def timeit(fn, flops, name):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"epi {name:34s}: {ms*1e3:7.1f} us  {flops/ms/1e9:7.1f} TFLOP/s", flush=True)
    x = torch.randn(M, 768, device=dev).to(torch.bfloat16)
    w1 = torch.randn(3072, 768, device=dev).to(torch.bfloat16)
    w2 = torch.randn(768, 3072, device=dev).to(torch.bfloat16)
    wq = torch.randn(2304, 768, device=dev).to(torch.bfloat16)
    b1 = torch.zeros(3072, device=dev); b2 = torch.zeros(768, device=dev); bq = torch.zeros(2304, device=dev)
    u = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16); pre = torch.empty_like(u)
    y = torch.empty(M, 768, device=dev, dtype=torch.bfloat16); res = torch.randn(M, 768, device=dev).to(torch.bfloat16)
    qkv = torch.empty(M, 2304, device=dev, dtype=torch.bfloat16)
    f1 = 2.0 * M * 3072 * 768
    timeit(lambda: K.gemm(x, w1, u, epilogue=K.EPI_GELU, bias=b1, aux_out=pre, p_drop=0.1, seed=1, site=2), f1, "fc1 GELU+drop (2 outputs)")
    timeit(lambda: K.gemm(x, w1, u, epilogue=K.EPI_GELU, bias=b1, aux_out=pre), f1, "fc1 GELU no drop")
    timeit(lambda: K.gemm(x, w1, u, bias=b1), f1, "fc1 plain bias")
    timeit(lambda: K.gemm(u, w2, y, epilogue=K.EPI_RES, bias=b2, aux_in=res, p_drop=0.1, seed=1, site=3), f1, "fc2 RES+drop")
    timeit(lambda: K.gemm(u, w2, y, epilogue=K.EPI_RES, bias=b2, aux_in=res), f1, "fc2 RES no drop")
    timeit(lambda: K.gemm(x, wq, qkv, bias=bq), 2.0 * M * 2304 * 768, "qkv plain bias")
    timeit(lambda: K.gemm(y, w2, u, b_mn=True, epilogue=K.EPI_DGELU, aux_in=pre, p_drop=0.1, seed=1, site=2), f1, "fc2 dgrad DGELU+drop")
    timeit(lambda: K.gemm(u, w1, y, b_mn=True, epilogue=K.EPI_ADD, aux_in=res), f1, "fc1 dgrad ADD")
    g1 = torch.zeros(3072, 768, device=dev); g2 = torch.zeros(768, 3072, device=dev)
    timeit(lambda: K.gemm(u, x, g1, a_mn=True, b_mn=True, epilogue=K.EPI_F32), f1, "fc1 wgrad f32")
    timeit(lambda: K.gemm(y, u, g2, a_mn=True, b_mn=True, epilogue=K.EPI_F32), f1, "fc2 wgrad f32")
    mk = torch.rand(3072, 768, device=dev) > 0.5
    timeit(lambda: K.gemm(u, x, g1, a_mn=True, b_mn=True, epilogue=K.EPI_F32, mask=mk), f1, "fc1 wgrad f32 masked")
if which in ("bn",):
    for (M, N, Kd) in [(24000, 2304, 768), (24000, 3072, 768), (24000, 768, 3072), (24000, 768, 768)]:
        a = torch.randn(M, Kd, device=dev).to(torch.bfloat16)
        b = torch.randn(N, Kd, device=dev).to(torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for bn in (128, 256):
            fn = lambda: K.gemm(a, b, out, block_n=bn)
            for _ in range(3): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"bn={bn} M={M} N={N} K={Kd}: {ms*1e3:.1f} us  {2*M*N*Kd/ms/1e9:.1f} TFLOP/s", flush=True)
