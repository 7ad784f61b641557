#!/bin/bash
# Every north_star configuration on one GPU (SURVEY 8d "How the five configs map to runs"): one JSON line each under
# gpurun_out/lines/.  usage (under gpurun): tools/bench_all.sh [steps]
S=${1:-20}
mkdir -p gpurun_out/lines
line() {  # name, bench.py args...
  local name=$1; shift
  timeout 400 python bench.py --steps $S --warmup 3 "$@" > gpurun_out/lines/$name.json 2> gpurun_out/lines/$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/lines/{n}.json").read().strip().splitlines()[-1])
    extra = ""
    if d.get("gpu_baseline"): extra += f"  stock-torch {d['gpu_baseline'].get('value', d['gpu_baseline'])}"
    if d.get("cpu_baseline"): extra += f"  cpu {d['cpu_baseline']['value']:.0f}"
    print(f"{n:28s} {d['value']/1e6:7.3f} M frames/s {d['ms_per_step']:8.3f} ms  e2e {d['e2e']['value']/1e6:7.3f} M  mfu {d['step_mfu']['frac']:.3f}  gemm {d['roofline']['frac']:.3f} attn {d['roofline_attention']['frac']:.3f}{extra}", flush=True)
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/lines/{n}.err").read()[-1200:])
PY
}
line cfg2_b32 
line cfg2_s1_b4x8 --batch 4 --accum 8 --no-cpu-baseline --no-gpu-baseline
line cfg2_b4 --batch 4 --no-cpu-baseline --no-gpu-baseline
line cfg3_10ms_h12 --mode head-pruning --frame 10 --heads 12 --batch 16 --no-cpu-baseline --no-gpu-baseline
line cfg3_10ms_h7 --mode head-pruning --frame 10 --heads 7 --batch 16 --no-cpu-baseline --no-gpu-baseline
line cfg3_10ms_h1 --mode head-pruning --frame 10 --heads 1 --batch 16 --no-cpu-baseline --no-gpu-baseline
line cfg4_row_weight --mode row+weight --no-cpu-baseline --no-gpu-baseline
line cfg4_weight --mode weight-pruning --no-cpu-baseline --no-gpu-baseline
line cfg4_row --mode row-pruning --no-cpu-baseline --no-gpu-baseline
line cfg5_distill --mode distillation --no-cpu-baseline --no-gpu-baseline
line extract_b32 --mode extract
line extract_e1 --mode extract --batch 2 --frames 791
