import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import synth_host_batch
from speech_ssl_compression_b200.trainer import TrainStep
from tools.bench_modes import build_expert
mode = sys.argv[1]
np.random.seed(0); torch.manual_seed(0)
B, T, D = 4, 256, 80
expert, _ = build_expert(mode, False, T)
expert.train()
ts = TrainStep(expert, B, T, D, use_graph=False)
hit = []
for n, p in expert.named_parameters():
    if p.requires_grad:
        p.register_hook(lambda g, n=n: hit.append((n, tuple(g.shape))))
f, l, p, lens = synth_host_batch(B, T, D, seed=1)
ts.load_batch(f, l, p, lens)
ts.run(); torch.cuda.synchronize()
print(mode, "leaf grads through autograd:", hit)
