"""Diagnostic: which autograd node runs on a non-capturing stream when a distillation step is captured."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import synth_host_batch
from speech_ssl_compression_b200 import ops
from speech_ssl_compression_b200.trainer import TrainStep
from tools.bench_modes import build_expert

mode = sys.argv[1] if len(sys.argv) > 1 else "distillation"
np.random.seed(0); torch.manual_seed(0)
B, T, D = 4, 256, 80
expert, _ = build_expert(mode, False, T)
expert.train()
ts = TrainStep(expert, B, T, D, use_graph=True)
f, l, p, lens = synth_host_batch(B, T, D, seed=1)
ts.load_batch(f, l, p, lens)

seen = []
for name in dir(ops):
    obj = getattr(ops, name)
    if isinstance(obj, type) and issubclass(obj, torch.autograd.Function) and obj is not torch.autograd.Function:
        orig = obj.backward
        def wrap(orig=orig, name=name):
            def bw(ctx, *a):
                st = torch.cuda.current_stream()
                cap = torch.cuda.is_current_stream_capturing()
                seen.append((name, st.cuda_stream, cap))
                return orig(ctx, *a)
            return staticmethod(bw)
        obj.backward = wrap()
try:
    ts.run()
    torch.cuda.synchronize()
    print("capture OK")
except Exception as e:
    print("capture FAILED:", str(e).splitlines()[0])
for s in seen[-40:]:
    print(s)
