"""Times the HBM-bound kernels at the bench shape and prints achieved GB/s (algorithmic bytes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_ssl_compression_b200 import kernels as K

dev = "cuda"
M, C = 24000, 768
x = torch.randn(M, C, device=dev).to(torch.bfloat16)
dy = torch.randn(M, C, device=dev).to(torch.bfloat16)
g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
y, mean, rstd = K.layernorm_fwd(x, g, b, 1e-5)
big = torch.randn(M, 3072, device=dev).to(torch.bfloat16)
out = torch.zeros(3072, device=dev)
w = torch.randn(3072, 768, device=dev); wb = torch.empty(3072, 768, device=dev, dtype=torch.bfloat16)


def t(name, fn, nbytes, reps=30):
    """`reps` calls captured into ONE CUDA graph (the Python wrappers need ~20 us of host time per call: eager timing of a
    15 us kernel measures the host), device-timed."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:28s}: {ms*1e3:7.1f} us  {nbytes/ms/1e6:7.0f} GB/s", flush=True)


t("ln_fwd", lambda: K.layernorm_fwd(x, g, b, 1e-5), 2 * M * C * 2)
t("ln_fwd + dropout", lambda: K.layernorm_fwd(x, g, b, 1e-5, p_drop=0.1, seed=1, site=1), 2 * M * C * 2)
t("ln_bwd", lambda: K.layernorm_bwd(dy, x, g, mean, rstd, dg, db), 3 * M * C * 2)
t("ln_bwd + dx_drop", lambda: K.layernorm_bwd(dy, x, g, mean, rstd, dg, db, want_drop=True, p_out=0.1, seed_out=1, site_out=2), 4 * M * C * 2)
t("colsum 3072", lambda: K.colsum_add(big, out), M * 3072 * 2)
t("colsum 768", lambda: K.colsum_add(x, out[:768]), M * 768 * 2)
t("weight_prep 3072x768", lambda: K.weight_prep(w, None, wb), 3072 * 768 * 6)
# practical ceilings at this size: plain copies of the same tensors (ATen vectorised copy kernel)
y2 = torch.empty_like(x)
t("copy 24000x768 bf16 (r+w)", lambda: y2.copy_(x), 2 * M * C * 2)
big2 = torch.empty_like(big)
t("copy 24000x3072 bf16 (r+w)", lambda: big2.copy_(big), 2 * M * 3072 * 2)
t("read-only sum 24000x3072 (torch.sum)", lambda: big.sum(), M * 3072 * 2)
