#!/usr/bin/env python
"""Benchmark of the MelHuBERT training hot path (BASELINE.json metric: train frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode MODE] [--frame 10|20]
                    [--batch B] [--accum A] [--heads H]

A "step" is one optimizer step (forward + backward + gradient all-reduce + clip + Adam) over
A micro-batches of B utterances x T frames of synthetic log-mel per GPU.  Default workload = cfg2 of
BASELINE.json: MelHuBERT-base 20 ms masked-prediction pre-training, bf16, B = 32, A = 1 (the
reference's 4 utterances x 8 gradient-accumulation micro-batches folded into one batch), T = 750,
dropout 0.1 as shipped.  The other north_star configurations are selected with --mode / --frame:
  cfg2 reference-faithful micro-batching (S1)   --batch 4 --accum 8
  cfg3 head-pruning, 10 ms                      --mode head-pruning --frame 10 --heads {12,7,1}
  cfg4 row pruning + 50 % unstructured masks    --mode row+weight
  cfg5 distillation 12L teacher -> 2L student   --mode distillation
  extraction (no_pred, get_hidden forward)      --mode extract
Under torchrun (N > 1) every rank runs the same per-GPU batch (weak scaling); the reported time is the max
over ranks.

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
    cfg = model_cfg(args.frame, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
    T, D = args.frames, 80 if args.frame == 20 else 40
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
           "config": dict(workload_config(args, per_gpu_batch=B), cuda_graph=False, same_config=False,
                          same_config_note="CPU arm: oracle port of the reference's pre-training step, fp32, dropout 0, bounded "
                                           "batch (its per-frame cost does not depend on B); GPU arm: bf16, dropout 0.1, the "
                                           "--batch given.  Dropout 0 favours the CPU arm (53 % of the reference's own CPU step "
                                           "is bernoulli_), so the per-frame ratio is conservative",
                          l2="CPU run: the working set streams through the host caches"),
           "cpu_baseline": {"value": val, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


MODE_CFG = {"pretrain": "cfg2", "weight-pruning": "cfg4 (masks only)", "row-pruning": "cfg4 (rows only)",
            "row+weight": "cfg4", "head-pruning": "cfg3", "distillation": "cfg5", "extract": "extraction"}


def workload_config(args, per_gpu_batch=None):
    B = per_gpu_batch if per_gpu_batch is not None else args.batch
    what = {
        "pretrain": "masked-prediction pre-training step",
        "weight-pruning": "pre-training step with 50 % global unstructured weight masks (dense-equivalent FLOPs)",
        "row-pruning": "pre-training step with FFN rows pruned to 1536",
        "row+weight": "pre-training step with FFN rows pruned to 1536 AND 50 % global unstructured weight masks",
        "head-pruning": f"pruning fine-tune step with {args.heads} of 12 attention heads kept per layer",
        "distillation": "distillation step, 12L teacher (no grad) -> 2L student, KD loss on all frames (nomasked, alpha 1)",
        "extract": "feature-extraction forward (eval, no_pred, get_hidden: 12 layer outputs)",
    }[args.mode]
    return {"workload": f"{MODE_CFG[args.mode]}: MelHuBERT-base (12L/768d/12h/3072ffn, 512 clusters) {args.frame} ms "
                        f"{what}, B={B} utterances x T={args.frames} frames x {args.accum} micro-batch(es) per GPU per step "
                        f"(reference: 4 utterances x 8 accumulation micro-batches per optimizer step)"
                        + ("" if args.mode == "extract" else ", dropout 0.1"),
            "mode": args.mode, "frame_ms": args.frame, "feat_dim": 80 if args.frame == 20 else 40,
            "per_gpu_batch": B, "frames": args.frames, "accum": args.accum, "global_batch": B * args.accum * args.gpus,
            "parallelism": f"dp{args.gpus}",
            "dp_transport": (os.environ.get("MH_DP_TRANSPORT", "peer") if args.gpus > 1 else None),
            "l2": ("bf16 weights (180 MB) + per-layer activations stream through HBM every forward: > 126 MB L2, no flush needed"
                   if args.mode == "extract" else
                   "every step streams ~1.6 GB of fp32 master / gradient / moment state (fused Adam) plus the activations "
                   "through HBM: >> 126 MB L2, no flush needed"),
            "cuda_graph": not args.no_graph}


def live_rooflines(run_once, peaks, B, T):
    """One eager (non-graph) pass of the workload with every tcgen05 GEMM launch and every fused-attention call
    bracketed by CUDA events on its own stream.  achieved = sum of algorithmic FLOPs / sum of bracket durations
    (GEMM: 2*M*N*K per launch; attention: 4*T*T*64 per (batch, head) forward, twice that backward -- the score
    recomputation of the backward is not counted).  Returns (gemm roofline, attention roofline).

    The host needs about as long to ENQUEUE an eager step (~320 launches + 2 event records per bracket) as the GPU
    needs to run it; whenever the GPU catches up with the host, the start event of the next bracket is stamped when
    the previous kernel ends but the kernel itself arrives microseconds later, and that idle gap would be counted
    as kernel time.  The stream is therefore held back by a spin kernel (torch.cuda._sleep, ~60 ms) while the host
    queues the whole pass: the kernels then run back to back exactly as inside the captured graph and each event
    pair brackets GPU time only."""
    import torch
    from speech_ssl_compression_b200 import kernels as K

    names = {K.EPI_BF16: "bias", K.EPI_GELU: "bias+GELU+dropout", K.EPI_RES: "bias+dropout+residual",
             K.EPI_F32: "wgrad fp32 reduce-add", K.EPI_DGELU: "dGELU+dropout", K.EPI_ADD: "+residual grad",
             K.EPI_DELTA: "dgrad + attention delta"}
    rec, arec = [], []
    real, real_af, real_ab = K.gemm, K.attn_fwd, K.attn_bwd

    def bracket():
        return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(a, b, out, *, a_mn=False, b_mn=False, epilogue=K.EPI_BF16, **kw):
        M, Kd = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
        N = b.shape[1] if b_mn else b.shape[0]
        e0, e1 = bracket()
        e0.record()
        r = real(a, b, out, a_mn=a_mn, b_mn=b_mn, epilogue=epilogue, **kw)
        e1.record()
        # algorithmic operand bytes: A + B read once, every output written once (bf16; fp32 for the wgrad reduce-add)
        nb = 2.0 * (M * Kd + N * Kd) + (4.0 if epilogue == K.EPI_F32 else 2.0) * M * N
        rec.append((epilogue, 2.0 * M * N * Kd, e0, e1, nb))
        return r

    def timed_af(qkv, kv_len, Bq, Tq, heads, **kw):
        e0, e1 = bracket()
        e0.record()
        r = real_af(qkv, kv_len, Bq, Tq, heads, **kw)
        e1.record()
        arec.append(("forward", 4.0 * Tq * Tq * 64 * Bq * heads, e0, e1))
        return r

    def timed_ab(qkv, kv_len, out, dout, lse, keep, Bq, Tq, heads, **kw):
        e0, e1 = bracket()
        e0.record()
        r = real_ab(qkv, kv_len, out, dout, lse, keep, Bq, Tq, heads, **kw)
        e1.record()
        arec.append(("backward", 8.0 * Tq * Tq * 64 * Bq * heads, e0, e1))
        return r

    K.gemm, K.attn_fwd, K.attn_bwd = timed, timed_af, timed_ab
    try:
        # how long the HOST needs to enqueue one bracketed pass (untimed rehearsal), then hold the stream for longer
        # than that while the measured pass is queued (see docstring)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_once()
        host_s = time.perf_counter() - t0
        torch.cuda.synchronize()
        rec.clear(); arec.clear()
        torch.cuda._sleep(int(max(0.06, 1.5 * host_s) * 1.9e9))
        run_once()
        torch.cuda.synchronize()
    finally:
        K.gemm, K.attn_fwd, K.attn_bwd = real, real_af, real_ab
    # what an event pair costs by itself: the same bracket around a one-thread kernel (mh_counter_add of 0), queued
    # behind the same kind of spin.  A bracket spans "previous work drained -> kernel launched -> kernel drained",
    # i.e. one un-overlapped launch that the kernel does not pay inside the captured graph.
    scratch = torch.zeros(1, device="cuda", dtype=torch.int64)
    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.005 * 1.9e9))
    cal = []
    for _ in range(32):
        c0, c1 = bracket()
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
        d = by.setdefault(names.get(epi, f"epilogue {epi}"), [0, 0.0, 0.0])
        d[0] += 1; d[1] += fl; d[2] += t
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    burst = peaks.get("bf16_tflops", 1590.0)
    peak_src = ("MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a running step)" if peaks
                else "fallback 1.4 PFLOP/s sustained")
    ach = flops / max(ms, 1e-9) / 1e9
    # DRAM traffic per launch of the same kernel family, from the committed ncu capture of this command
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "gemm_dram_traffic.json")) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except (OSError, KeyError, ValueError):
        pass
    gemm = {"bound": "tensor", "kernel": "mh::gemm_pair_kernel<EPI, A_MN, B_MN> / mh::gemm_kernel<BN, ...> (all tcgen05 GEMM launches of one step)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
            "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, average over the family)",
            "traffic_source": traffic_src, "algorithmic_bytes_per_launch": abytes / max(len(rec), 1),
            "peak_source": peak_src,
            "launches": len(rec), "timing": "CUDA-event pair per launch on the launching stream, step pre-queued behind a spin kernel",
            "event_bracket_us": bracket_us,
            "achieved_net_of_bracket": flops / max(ms - len(rec) * bracket_us * 1e-3, 1e-6) / 1e9,
            "bracket_note": "event_bracket_us = median event-pair time around a one-thread kernel; achieved / frac are RAW "
                            "(bracket included, conservative); achieved_net_of_bracket subtracts it per launch and is what "
                            "the in-graph CUPTI timeline shows (profiles/*timeline_graph_step*)",
            "gemm_ms_per_step": ms, "gemm_tflop_per_step": flops / 1e12,
            "by_epilogue": {k: {"launches": v[0], "tflops": v[1] / v[2] / 1e9, "ms": v[2]} for k, v in by.items()},
            "frac_of_burst_peak": ach / burst}
    aby = {}
    aflops = ams = 0.0
    for kind, fl, e0, e1 in arec:
        t = e0.elapsed_time(e1)
        aflops += fl
        ams += t
        d = aby.setdefault(kind, [0, 0.0, 0.0])
        d[0] += 1; d[1] += fl; d[2] += t
    aach = aflops / max(ams, 1e-9) / 1e9
    attn = {"bound": "tensor", "kernel": "mh::attn_fwd2_kernel / mh::attn_bwd_kernel (+ attn_delta, dq_finish, dQ workspace memset: "
                                         "everything one mh_attn_fwd / mh_attn_bwd call launches)",
            "achieved": aach, "peak": peak, "unit": "TFLOP/s", "frac": aach / peak, "traffic": None, "peak_source": peak_src,
            "calls": len(arec), "timing": "CUDA-event pair per C-ABI call on the launching stream, pass pre-queued behind a spin kernel",
            "attention_ms_per_step": ams, "attention_tflop_per_step": aflops / 1e12,
            "by_pass": {k: {"calls": v[0], "tflops": v[1] / v[2] / 1e9, "us_per_call": v[2] / v[0] * 1e3} for k, v in aby.items()},
            "frac_of_burst_peak": aach / burst,
            "gemm_plus_attention": {"tflop": (flops + aflops) / 1e12, "ms": ms + ams,
                                    "tflops": (flops + aflops) / max(ms + ams, 1e-9) / 1e9,
                                    "frac_sustained": (flops + aflops) / max(ms + ams, 1e-9) / 1e9 / peak,
                                    "frac_burst": (flops + aflops) / max(ms + ams, 1e-9) / 1e9 / burst}}
    return gemm, attn


# ------------------------------------------------------------------------------------------------
def gpu_stock_torch_baseline(T, D, frame, steps=3):
    """Informational "practical bar" (SURVEY 8d last row): the reference algorithm as stock PyTorch ops on the same
    B200 -- the oracle port moved to CUDA under bf16 autocast (the reference runs fp16 autocast, runner.py:363), B = 4
    utterances (its shipped micro-batch), dropout 0, forward + backward + clip + Adam, CUDA-event timed."""
    import numpy as np
    import torch
    from oracle import melhubert_oracle as O

    cfg = model_cfg(frame, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
    B = 4
    sd = {k: v.cuda().requires_grad_(True) for k, v in O.synth_state_dict(cfg, seed=7).items()}
    opt = torch.optim.Adam(list(sd.values()), lr=1e-4)
    lens = [T] * B
    feat, label, pad = (x.cuda() for x in O.synth_batch(B, T, D, lens))
    np.random.seed(1337)
    mask = torch.from_numpy(O.span_mask(B, T, lens, cfg["mask_prob"], cfg["mask_length"])).cuda()

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = O.model_forward(sd, cfg, feat, pad, label, mask_indices=mask)
            loss = O.ce_mean(out["logit_m"].float(), out["label_m"])
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(sd.values()), 10.0)
        opt.step()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": B * T / (ms / 1e3), "unit": "frames/s", "ms_per_step": ms, "kind": "oracle port on cuda:0, torch ops under bf16 autocast",
            "sample": f"{steps} steps of {B} utterances x {T} frames, fwd+bwd+clip+Adam, dropout 0 (cuBLAS / cuDNN / ATen kernels)",
            "note": "informational: what stock PyTorch gives the reference algorithm on this GPU; not the reference arm"}


def bench_extract(args, rank, world, local, peaks):
    """--mode extract: frames/s of the ``no_pred=True, get_hidden=True`` eval forward (extract_feature.py:145-146)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from speech_ssl_compression_b200 import kernels as K
    from speech_ssl_compression_b200.model import MelHuBERTConfig, MelHuBERTModel

    B, T, D = args.batch, args.frames, 80 if args.frame == 20 else 40
    torch.manual_seed(1337)
    model = MelHuBERTModel(MelHuBERTConfig(model_cfg(args.frame))).cuda().eval()
    batches = [synth_host_batch(B, T, D, seed=2024 + 97 * rank + i) for i in range(4)]
    feat, pad = torch.zeros(B, T, D, device="cuda"), torch.ones(B, T, device="cuda")
    h_out = torch.empty(B, T, 768).pin_memory()
    state = {}

    def fwd(lens):
        with torch.no_grad():
            out = model(feat, pad, get_hidden=True, no_pred=True, valid_lens=lens)
        state["out"] = out
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def load(i):
        f, _, p, lens = batches[i % len(batches)]
        feat.copy_(f, non_blocking=True)
        pad.copy_(p, non_blocking=True)
        return lens

    for i in range(args.warmup):
        fwd(load(i))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lens = load(0)
    barrier()
    n0 = K.L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fwd(lens)
    e1.record()
    barrier()
    ms_resident = e0.elapsed_time(e1) / args.steps
    launches = K.L.launch_count() - n0
    # end to end: pinned features in, last-layer features out (what extract_feature.py returns), every step
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = fwd(load(i))
        h_out.copy_(out[0].view(B, T, 768), non_blocking=True)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_resident, ms_e2e], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_resident, ms_e2e = float(t[0]), float(t[1])
    roof, roof_attn = live_rooflines(lambda: fwd(lens), peaks, B, T)
    if rank != 0:
        return None
    frames = B * T * world
    fl = fwd_flops_per_frame(T, D, 12, 12, 3072, 0.0) / 1e6
    sustained = peaks.get("bf16_tflops_sustained", 1400.0)
    return {
        "metric": "extraction frames/s", "value": frames / (ms_resident / 1e3), "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_resident, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": dict(workload_config(args), cuda_graph=False),
        "e2e": {"value": frames / (ms_e2e / 1e3), "unit": "frames/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": B * T * (D + 1) * 4, "d2h_bytes_per_step": B * T * 768 * 4,
                "pipeline": "pinned features + pad mask copied in, last-layer features (fp32) copied out, every forward"},
        "gpu_launches": int(launches), "launches_per_step": int(launches // max(args.steps, 1)), "clocks": clocks,
        "roofline": roof, "roofline_attention": roof_attn,
        "step_mfu": {"algorithmic_tflops": fl * 1e6 * frames / (ms_resident / 1e3) / 1e12 / world,
                     "peak_sustained_tflops": sustained,
                     "frac": fl * 1e6 * frames / (ms_resident / 1e3) / 1e12 / world / sustained, "mflop_per_frame_fwd": fl},
        "cpu_baseline": None,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="pretrain", choices=list(MODE_CFG))
    ap.add_argument("--frame", type=int, default=20, choices=[10, 20], help="frame period in ms (10: D_in 40, T 1500, mask spans of 10)")
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU per micro-batch")
    ap.add_argument("--accum", type=int, default=1, help="gradient-accumulation micro-batches per optimizer step")
    ap.add_argument("--heads", type=int, default=6, help="--mode head-pruning: attention heads kept per layer")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.frames is None:
        args.frames = 750 if args.frame == 20 else 1500

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
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if args.mode == "extract":
        out = bench_extract(args, rank, world, local, peaks)
        if out is not None:
            print(json.dumps(out), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    B, T, D, A = args.batch, args.frames, 80 if args.frame == 20 else 40, max(args.accum, 1)
    expert, flops_train_per_frame = build_expert(args.mode, world > 1, T, frame=args.frame, heads=args.heads)
    expert.train()
    ts = TrainStep(expert, B, T, D, lr=1e-4, max_norm=10.0, use_graph=not args.no_graph, accum=A)
    batches = [synth_host_batch(B, T, D, seed=2024 + 97 * rank + i) for i in range(max(4, A))]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (includes CUDA-graph capture)
    for i in range(args.warmup):
        for a in range(A):
            ts.load_batch(*batches[(i * A + a) % len(batches)])
            ts.run()
        loss0 = ts.read_loss()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- (1) device-resident: inputs already in HBM, K steps (of A micro-batches) back to back
    ts.load_batch(*batches[0])
    barrier()
    n0 = K.L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = []
    e0.record()
    for _ in range(args.steps):
        for a in range(A):
            ts.run()
        m = torch.cuda.Event(enable_timing=True)
        m.record()  # (a time stamp between two graph launches: no synchronisation, no extra work)
        marks.append(m)
    e1.record()
    barrier()
    ms_resident = e0.elapsed_time(e1) / args.steps
    series = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
    eager_launches = K.L.launch_count() - n0

    # ---- (2) end to end through the public step API: H2D of every micro-batch (pinned) + step + D2H of the loss
    #      Every micro-batch comes from pinned host memory and every step's loss goes back to the host; the input
    #      pipeline is the one the training loop uses (runner.py: TrainStep.stage_batch / commit_staged /
    #      read_loss_async): micro-batch i+1 is drawn (host span masks) and copied on a copy stream while micro-batch i
    #      runs, the loss of step i is collected after step i+1 has been launched.
    barrier()
    t0 = time.perf_counter()
    ts.stage_batch(*batches[0])
    pending = None
    n_micro = args.steps * A
    for i in range(n_micro):
        ts.commit_staged()
        last = ts.run()
        if last:
            h = ts.read_loss_async()
        if i + 1 < n_micro:
            ts.stage_batch(*batches[(i + 1) % len(batches)])
        if last:
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

    # ---- rooflines of the two tensor-core kernel families (tcgen05 GEMMs, fused attention), measured LIVE inside
    #      one extra eager optimizer step: a CUDA-event pair around every launch / call on the launching stream
    def one_eager_step():
        was = ts.use_graph
        ts.use_graph = False
        try:
            for a in range(A):
                ts.run()
        finally:
            ts.use_graph = was

    ts.load_batch(*batches[0])
    roof, roof_attn = live_rooflines(one_eager_step, peaks, B, T)  # every rank runs it: the step contains collectives
    if rank != 0:
        dist.barrier()
        if not ts.use_graph:
            dist.destroy_process_group()
        return
    frames = B * T * A * world
    value = frames / (ms_resident / 1e3)
    step_flops = flops_train_per_frame * 1e6 * frames
    sustained = peaks.get("bf16_tflops_sustained", 1400.0)
    out = {
        "metric": "train frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_resident, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": frames / (ms_e2e / 1e3), "unit": "frames/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": ts.h2d_bytes * A, "d2h_bytes_per_step": ts.d2h_bytes,
                "pipeline": "micro-batch i+1 (host span masks + pinned H2D on a copy stream) staged while micro-batch i runs; "
                            "loss of step i collected after step i+1 is launched"},
        "gpu_launches": int(ts.launches_per_step * args.steps if ts.use_graph else eager_launches),
        "launches_per_step": int(ts.launches_per_step),
        "clocks": clocks,
        "ms_per_step_series": [round(x, 3) for x in series],  # rank 0; a rising series = the power cap pulling clocks down
        "roofline": roof,
        "roofline_attention": roof_attn,
        "step_mfu": {"algorithmic_tflops": step_flops / (ms_resident / 1e3) / 1e12 / world,
                     "peak_sustained_tflops": sustained,
                     "frac": step_flops / (ms_resident / 1e3) / 1e12 / world / sustained,
                     "frac_of_burst_peak": step_flops / (ms_resident / 1e3) / 1e12 / world / peaks.get("bf16_tflops", 1590.0),
                     "mflop_per_frame_train": flops_train_per_frame},
        "loss": loss,
    }
    if not args.no_cpu_baseline and world == 1:
        cfg = model_cfg(args.frame, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
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
                                         f"(oracle port of the reference's pre-training step; per-frame cost of the CPU path "
                                         f"does not depend on the batch, so the comparison is per frame)"}
    else:
        out["cpu_baseline"] = None
    if not args.no_gpu_baseline and world == 1:
        del ts, expert
        torch.cuda.empty_cache()
        try:
            out["gpu_baseline"] = gpu_stock_torch_baseline(T, D, args.frame)
        except Exception as e:  # informational only
            out["gpu_baseline"] = {"error": str(e)[:200]}
        print(json.dumps(out), flush=True)
        return
    out["config"]["cuda_graph"] = bool(ts.use_graph)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        if not ts.use_graph:  # a captured graph still references the communicator: let process exit release it
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
