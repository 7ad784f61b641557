#!/bin/bash
# One N-GPU box: DP equivalence test, then bench lines (pretrain peer / nccl, distillation peer) -> gpurun_out/scale_*.json
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "peer and not sm" 2>&1 | tail -3
run() {  # tag transport mode
  MH_DP_TRANSPORT=$2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $N --steps 30 --warmup 3 --mode $3 > gpurun_out/scale_$1.json 2> gpurun_out/scale_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/scale_{tag}.json").read().strip().splitlines()[-1])
    s = d["ms_per_step_series"]
    print(f"{tag:22s} n={d['n_gpus']} {d['value']/1e6:7.3f} M frames/s  {d['ms_per_step']:7.3f} ms  e2e {d['e2e']['ms_per_step']:7.3f} ms  first5 {sum(s[:5])/5:6.2f} last5 {sum(s[-5:])/5:6.2f}", flush=True)
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/scale_{tag}.err").read()[-1500:])
PY
}
one() {  # tag mode
  timeout 300 python bench.py --steps 30 --warmup 3 --mode $2 --no-cpu-baseline --no-gpu-baseline > gpurun_out/scale_$1.json 2> gpurun_out/scale_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
d = json.loads(open(f"gpurun_out/scale_{tag}.json").read().strip().splitlines()[-1])
print(f"{tag:22s} n=1 {d['value']/1e6:7.3f} M frames/s  {d['ms_per_step']:7.3f} ms", flush=True)
PY
}
one pretrain_n1 pretrain
run pretrain_peer_n$N peer pretrain
run pretrain_nccl_n$N nccl pretrain
one distill_n1 distillation
run distill_peer_n$N peer distillation
