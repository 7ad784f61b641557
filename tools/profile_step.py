"""In-stream kernel timeline of one optimizer step (torch.profiler / CUPTI): per-kernel totals inside
the running step (warm caches, real clocks, real overlap), GPU busy time vs wall time.  Complements the
ncu launch list (cold-cache, serialised).  Usage: python tools/profile_step.py [--mode pretrain] [--graph]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from bench import synth_host_batch
from speech_ssl_compression_b200.trainer import TrainStep
from tools.bench_modes import build_expert

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="pretrain")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--frames", type=int, default=750)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--list", type=int, default=0, help="also print the last N kernel launches in stream order (name, us)")
args = ap.parse_args()

np.random.seed(1337)
torch.manual_seed(1337)
B, T, D = args.batch, args.frames, 80
expert, _ = build_expert(args.mode, False, T)
expert.train()
ts = TrainStep(expert, B, T, D, lr=1e-4, max_norm=10.0, use_graph=args.graph)
f, l, p, lens = synth_host_batch(B, T, D, seed=2024)
for _ in range(4):
    ts.load_batch(f, l, p, lens)
    ts.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    e0.record()
    for _ in range(args.steps):
        ts.run()
    e1.record()
    torch.cuda.synchronize()
wall_ms = e0.elapsed_time(e1) / args.steps
agg = collections.defaultdict(lambda: [0, 0.0])
spans = []
order = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0][:90]
        order.append((ev.time_range.start, name, ev.time_range.end - ev.time_range.start))
        agg[name][0] += 1
        agg[name][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        spans.append((ev.time_range.start, ev.time_range.end))
tot = sum(v[1] for v in agg.values())
spans.sort()
busy, cur_s, cur_e = 0.0, None, None
for s, e in spans:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
if cur_e is not None:
    busy += cur_e - cur_s
print(f"wall {wall_ms:.2f} ms/step   sum of kernel durations {tot / 1e3 / args.steps:.2f} ms/step   "
      f"GPU busy (union) {busy / 1e3 / args.steps:.2f} ms/step")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    print(f"{t / args.steps:10.1f} us/step {100 * t / tot:5.1f}%  n/step={n / args.steps:6.1f} avg={t / n:8.1f} us  {k}")
if args.list:
    order.sort()
    print(f"--- last {args.list} launches in stream order")
    for _, name, us in order[-args.list:]:
        print(f"{us:9.1f} us  {name}")
