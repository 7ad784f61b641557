#!/usr/bin/env python
"""Benchmark of the MelHuBERT training hot path (BASELINE.json metric: train frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode MODE]

A "step" is one optimizer step (forward + backward + gradient all-reduce + clip + Adam) over
one synthetic log-mel batch of B utterances x T frames per GPU.  Default workload = cfg2 of
BASELINE.json: MelHuBERT-base 20 ms masked-prediction pre-training, bf16, B = 32 (the
reference's 4 utterances x 8 gradient-accumulation micro-batches = one optimizer step),
T = 750, dropout 0.1 as shipped.  Under torchrun (N > 1) every rank runs the same per-GPU
batch (weak scaling); the reported time is the max over ranks.

Rank 0 prints ONE JSON line (see DESIGN.md "Measurement" for every key).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_MFLOP_PER_FRAME = {  # F_fwd(T) per frame, dense-equivalent
    "pretrain": 207.5, "weight-pruning": 207.5, "distillation": 336.1 / 1.0,
}


def model_cfg(frame=20, layers=12, ffn=3072, **over):
    cfg = dict(feat_emb_dim=80 if frame == 20 else 40, pos_emb_type="conv", pos_conv_depth=1, conv_pos=128,
               conv_pos_groups=16, encoder_layers=layers, encoder_embed_dim=768, encoder_ffn_embed_dim=ffn,
               encoder_attention_heads=12, activation_fn="gelu", layer_norm_first=False, attention_type="original",
               num_cluster=512, pred_masked_weight=1.0, pred_nomask_weight=0.0, skip_masked=False, skip_nomask=True,
               mask_prob=0.7, mask_length=5 if frame == 20 else 10, mask_selection="static", mask_other=0.0,
               no_mask_overlap=False, mask_min_space=1, learnable_mask_emb=False, mask_before_proj=True, dropout=0.1,
               attention_dropout=0.1, activation_dropout=0.1, encoder_layerdrop=0.0)
    cfg.update(over)
    return cfg


def fwd_flops_per_frame(T, d_in, layers, heads, ffn, rho, d=768, k=512):
    """2*D_in*d + 2*d*(d/16)*128 + sum_l [8 d e + 4 d f + 4 T e] + rho*2*d*K"""
    e = 64 * heads
    return 2 * d_in * d + 2 * d * (d // 16) * 128 + layers * (8 * d * e + 4 * d * ffn + 4 * T * e) + rho * 2 * d * k


def synth_host_batch(B, T, D, seed, ragged=True):
    import torch

    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(B, T, D, generator=g)
    label = torch.randint(0, 512, (B, T), generator=g)
    # length-sorted bucket like the reference's bucketing dataset: longest first, >= 80 % of T
    lens = sorted([T - int(x) for x in torch.randint(0, T // 5, (B,), generator=g)], reverse=True) if ragged else [T] * B
    lens[0] = T
    pad = torch.ones(B, T)
    for i, l in enumerate(lens):
        pad[i, l:] = 0
        label[i, l:] = -100
        feat[i, l:] = 0
    return feat.pin_memory(), label.pin_memory(), pad.pin_memory(), lens


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            try:
                power.append(float(parts[2]))
            except ValueError:
                pass
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        busy = [x for x in sm if smax and x > 0.5 * smax] or sm
        med = busy[len(busy) // 2] if busy else None
        power.sort()
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": power[-1] if power else None}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's training step (fp32, all host threads)
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(cfg, B, T, D, mode="pretrain"):
    """Returns (step_fn, frames_per_step).  step_fn() runs forward + backward + Adam of the
    reference algorithm (oracle restatement, dropout 0 -- see BASELINE.md §4) on CPU."""
    import numpy as np
    import torch
    from oracle import melhubert_oracle as O

    torch.set_num_threads(os.cpu_count())
    c = dict(cfg)
    sd = {k: v.clone().requires_grad_(True) for k, v in O.synth_state_dict(c, seed=7).items()}
    opt = torch.optim.Adam(list(sd.values()), lr=1e-4)
    lens = [T] * B
    feat, label, pad = O.synth_batch(B, T, D, lens)

    def step():
        np.random.seed(1337)
        mask = torch.from_numpy(O.span_mask(B, T, lens, c["mask_prob"], c["mask_length"]))
        out = O.model_forward(sd, c, feat, pad, label, mask_indices=mask)
        loss = O.ce_mean(out["logit_m"], out["label_m"])
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(sd.values()), 10.0)
        opt.step()
        return float(loss)

    return step, B * T


def run_reference_arm(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; the Python reference
    itself cannot travel to the GPU box) on the host cores, same metric / config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = model_cfg(dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
    T, D = args.frames, 80
    # bounded sample: probe one utterance, then size B so that (K + W) steps stay within ~150 s
    step, frames = cpu_reference_step_fn(cfg, 1, T, D)
    t0 = time.time(); step(); probe = time.time() - t0
    budget = 150.0 / max(1, args.steps + args.warmup)
    B = max(1, min(4, int(budget / max(probe, 1e-3))))
    if B > 1:
        step, frames = cpu_reference_step_fn(cfg, B, T, D)
    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = (time.time() - t0) / args.steps
    val = frames / dt
    sample = f"{B} utterance(s) x {T} frames per step, fwd+bwd+clip+Adam, fp32, dropout 0, oracle port of the reference"
    out = {"impl": "reference", "metric": "train frames/s", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": dict(workload_config(args, per_gpu_batch=B), cuda_graph=False,
                          l2="CPU run: the working set streams through the host caches"),
           "cpu_baseline": {"value": val, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def workload_config(args, per_gpu_batch=None):
    B = per_gpu_batch if per_gpu_batch is not None else args.batch
    return {"workload": f"cfg2: MelHuBERT-base (12L/768d/12h/3072ffn, 512 clusters) 20 ms masked-prediction "
                        f"pre-training step, {args.mode}, B={B} utterances x T={args.frames} frames per GPU "
                        f"(reference: 4 utterances x 8 accumulation micro-batches per optimizer step), dropout 0.1",
            "per_gpu_batch": B, "frames": args.frames, "global_batch": B * args.gpus, "parallelism": f"dp{args.gpus}",
            "l2": "per-step working set (several GB of activations) >> 126 MB L2, no flush needed",
            "cuda_graph": not args.no_graph}


def live_gemm_roofline(ts, batch, peaks):
    """One eager (non-graph) optimizer step with every tcgen05 GEMM launch bracketed by CUDA events on its
    own stream.  achieved = sum of algorithmic FLOPs (2*M*N*K per launch) / sum of launch durations.

    The host needs about as long to ENQUEUE an eager step (~320 launches + 2 event records per GEMM) as the GPU
    needs to run it; whenever the GPU catches up with the host, the start event of the next GEMM is stamped when
    the previous kernel ends but the GEMM itself arrives microseconds later, and that idle gap was counted as GEMM
    time (the first bench lines of this round under-reported the family by ~25 % against the in-graph timeline,
    profiles/r01_h_timeline_graph_step.txt).  The stream is therefore held back by a spin kernel
    (torch.cuda._sleep, ~60 ms) while the host queues the whole step: the kernels then run back to back exactly as
    inside the captured graph and each event pair brackets GPU time only."""
    import torch
    from speech_ssl_compression_b200 import kernels as K
    from speech_ssl_compression_b200 import ops

    names = {K.EPI_BF16: "bias", K.EPI_GELU: "bias+GELU+dropout", K.EPI_RES: "bias+dropout+residual",
             K.EPI_F32: "wgrad fp32 reduce-add", K.EPI_DGELU: "dGELU+dropout", K.EPI_ADD: "+residual grad"}
    rec = []
    real = K.gemm

    def timed(a, b, out, *, a_mn=False, b_mn=False, epilogue=K.EPI_BF16, **kw):
        M, Kd = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
        N = b.shape[1] if b_mn else b.shape[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = real(a, b, out, a_mn=a_mn, b_mn=b_mn, epilogue=epilogue, **kw)
        e1.record()
        # algorithmic operand bytes: A + B read once, every output written once (bf16; fp32 for the wgrad reduce-add)
        nb = 2.0 * (M * Kd + N * Kd) + (4.0 if epilogue == K.EPI_F32 else 2.0) * M * N
        rec.append((epilogue, 2.0 * M * N * Kd, e0, e1, nb))
        return r

    f, l, p, lens = batch
    was_graph = ts.use_graph
    ts.use_graph = False
    K.gemm = timed
    ops.K.gemm = timed
    try:
        ts.load_batch(f, l, p, lens)
        torch.cuda.synchronize()
        torch.cuda._sleep(int(0.06 * 1.9e9))  # hold the stream while the host enqueues the step (see docstring)
        ts.run()
        torch.cuda.synchronize()
    finally:
        K.gemm = real
        ops.K.gemm = real
        ts.use_graph = was_graph
    # what an event pair costs by itself: the same bracket around a one-thread kernel (mh_counter_add of 0), queued
    # behind the same kind of spin.  A bracket spans "previous work drained -> kernel launched -> kernel drained",
    # i.e. one un-overlapped launch that the kernel does not pay inside the captured graph.
    scratch = torch.zeros(1, device="cuda", dtype=torch.int64)
    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.005 * 1.9e9))
    cal = []
    for _ in range(32):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(); K.counter_add(scratch, 0); c1.record()
        cal.append((c0, c1))
    torch.cuda.synchronize()
    cal_us = sorted(a.elapsed_time(b) * 1e3 for a, b in cal)
    bracket_us = cal_us[len(cal_us) // 2]
    by = {}
    flops = ms = 0.0
    abytes = 0.0
    for epi, fl, e0, e1, nb in rec:
        t = e0.elapsed_time(e1)
        flops += fl
        ms += t
        abytes += nb
        d = by.setdefault(names[epi], [0, 0.0, 0.0])
        d[0] += 1; d[1] += fl; d[2] += t
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    ach = flops / ms / 1e9
    # DRAM traffic per launch of the same kernel family, from the committed ncu capture of this command
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "gemm_dram_traffic.json")) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except (OSError, KeyError, ValueError):
        pass
    return {"bound": "tensor", "kernel": "mh::gemm_pair_kernel<EPI, A_MN, B_MN> / mh::gemm_kernel<BN, ...> (all tcgen05 GEMM launches of one optimizer step)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
            "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, average over the family)",
            "traffic_source": traffic_src, "algorithmic_bytes_per_launch": abytes / max(len(rec), 1),
            "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a running step)" if peaks
                            else "fallback 1.4 PFLOP/s sustained"),
            "launches": len(rec), "timing": "CUDA-event pair per launch on the launching stream, step pre-queued behind a spin kernel",
            "event_bracket_us": bracket_us,
            "achieved_net_of_bracket": flops / max(ms - len(rec) * bracket_us * 1e-3, 1e-6) / 1e9,
            "bracket_note": "event_bracket_us = median event-pair time around a one-thread kernel; achieved / frac are RAW "
                            "(bracket included, conservative); achieved_net_of_bracket subtracts it per launch and is what "
                            "the in-graph CUPTI timeline shows (profiles/*timeline_graph_step*)",
            "gemm_ms_per_step": ms, "gemm_tflop_per_step": flops / 1e12,
            "by_epilogue": {k: {"launches": v[0], "tflops": v[1] / v[2] / 1e9, "ms": v[2]} for k, v in by.items()},
            "frac_of_burst_peak": ach / peaks.get("bf16_tflops", 1590.0)}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="pretrain", choices=["pretrain", "distillation", "weight-pruning", "head-pruning", "row-pruning"])
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU per optimizer step")
    ap.add_argument("--frames", type=int, default=750)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from speech_ssl_compression_b200 import kernels as K
    from speech_ssl_compression_b200.parallel import init_distributed
    from speech_ssl_compression_b200.trainer import TrainStep
    from tools.bench_modes import build_expert

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    rank, world = init_distributed("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    np.random.seed(1337 + rank)
    torch.manual_seed(1337)
    B, T, D = args.batch, args.frames, 80
    expert, flops_train_per_frame = build_expert(args.mode, world > 1, T)
    expert.train()
    ts = TrainStep(expert, B, T, D, lr=1e-4, max_norm=10.0, use_graph=not args.no_graph)
    batches = [synth_host_batch(B, T, D, seed=2024 + 97 * rank + i) for i in range(4)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (includes CUDA-graph capture)
    for i in range(args.warmup):
        f, l, p, lens = batches[i % len(batches)]
        ts.load_batch(f, l, p, lens)
        ts.run()
        loss0 = ts.read_loss()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- (1) device-resident: inputs already in HBM, K steps back to back
    f, l, p, lens = batches[0]
    ts.load_batch(f, l, p, lens)
    barrier()
    n0 = K.L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = []
    e0.record()
    for _ in range(args.steps):
        ts.run()
        m = torch.cuda.Event(enable_timing=True)
        m.record()  # (a time stamp between two graph launches: no synchronisation, no extra work)
        marks.append(m)
    e1.record()
    barrier()
    ms_resident = e0.elapsed_time(e1) / args.steps
    series = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
    eager_launches = K.L.launch_count() - n0

    # ---- (2) end to end through the public step API: H2D of the batch (pinned) + step + D2H of the loss
    #      Every step's batch comes from pinned host memory and every step's loss goes back to the host; the input
    #      pipeline is the one a training loop uses (trainer.TrainStep.stage_batch / commit_staged / read_loss_async):
    #      batch i+1 is drawn (host span masks) and copied on a copy stream while step i runs, the loss of step i is
    #      collected after step i+1 has been launched.
    barrier()
    t0 = time.perf_counter()
    ts.stage_batch(*batches[0])
    pending = None
    for i in range(args.steps):
        ts.commit_staged()
        ts.run()
        h = ts.read_loss_async()
        if i + 1 < args.steps:
            ts.stage_batch(*batches[(i + 1) % len(batches)])
        if pending is not None:
            loss = ts.collect_loss(pending)
        pending = h
    loss = ts.collect_loss(pending)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms_resident, ms_e2e], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_resident, ms_e2e = float(t[0]), float(t[1])

    # ---- roofline of the dominant kernel family (the tcgen05 GEMM: ~45 % of the step), measured LIVE inside
    #      one extra eager optimizer step: a CUDA-event pair around every mh_gemm launch on the launching stream
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roof = live_gemm_roofline(ts, batches[0], peaks)  # every rank runs it: the step contains collectives
    if rank != 0:
        dist.barrier()
        if not ts.use_graph:
            dist.destroy_process_group()
        return
    frames = B * T * world
    value = frames / (ms_resident / 1e3)
    step_flops = flops_train_per_frame * 1e6 * frames
    sustained = peaks.get("bf16_tflops_sustained", 1400.0)
    out = {
        "metric": "train frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_resident, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": frames / (ms_e2e / 1e3), "unit": "frames/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": ts.h2d_bytes, "d2h_bytes_per_step": ts.d2h_bytes,
                "pipeline": "batch i+1 (host span masks + pinned H2D on a copy stream) staged while step i runs; "
                            "loss of step i collected after step i+1 is launched"},
        "gpu_launches": int(ts.launches_per_step * args.steps if ts.use_graph else eager_launches),
        "launches_per_step": int(ts.launches_per_step),
        "clocks": clocks,
        "ms_per_step_series": [round(x, 3) for x in series],  # rank 0; a rising series = the power cap pulling clocks down
        "roofline": roof,
        "step_mfu": {"algorithmic_tflops": step_flops / (ms_resident / 1e3) / 1e12 / world,
                     "peak_sustained_tflops": sustained,
                     "frac": step_flops / (ms_resident / 1e3) / 1e12 / world / sustained,
                     "mflop_per_frame_train": flops_train_per_frame},
        "loss": loss,
    }
    if not args.no_cpu_baseline and world == 1:
        cfg = model_cfg(dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
        step, fr = cpu_reference_step_fn(cfg, 1, T, D)
        step()
        t0 = time.time()
        n = 0
        while time.time() - t0 < 12.0 or n < 1:
            step()
            n += 1
        dt = (time.time() - t0) / n
        out["cpu_baseline"] = {"value": fr / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"{n} step(s) of 1 utterance x {T} frames, fwd+bwd+clip+Adam, fp32, dropout 0 "
                                         f"(oracle port of the reference training step)"}
    else:
        out["cpu_baseline"] = None
    out["config"]["cuda_graph"] = bool(ts.use_graph)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        if not ts.use_graph:  # a captured graph still references the communicator: let process exit release it
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
