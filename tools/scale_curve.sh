#!/bin/bash
# Same-box weak-scaling curve (run under gpurun --gpus 8): N = 1, 2, 4, 8 for pre-training and distillation, default
# (peer copy-engine) transport -> gpurun_out/curve_<mode>_n<N>.json
mkdir -p gpurun_out
for mode in pretrain distillation; do
  for N in 1 2 4 8; do
    if [ $N = 1 ]; then
      timeout 300 python bench.py --steps 20 --warmup 3 --mode $mode --no-cpu-baseline --no-gpu-baseline > gpurun_out/curve_${mode}_n$N.json 2> gpurun_out/curve_${mode}_n$N.err
    else
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N \
        bench.py --gpus $N --steps 20 --warmup 3 --mode $mode > gpurun_out/curve_${mode}_n$N.json 2> gpurun_out/curve_${mode}_n$N.err
    fi
    python - "$mode" "$N" <<'PY'
import json, sys
mode, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/curve_{mode}_n{n}.json").read().strip().splitlines()[-1])
    print(f"{mode:13s} n={d['n_gpus']} {d['value']/1e6:7.3f} M frames/s  {d['ms_per_step']:7.3f} ms  e2e {d['e2e']['value']/1e6:7.3f} M  gemm {d['roofline']['achieved']:.0f} TFLOP/s", flush=True)
except Exception as e:
    print(mode, n, "FAILED", e); print(open(f"gpurun_out/curve_{mode}_n{n}.err").read()[-1200:])
PY
  done
done
