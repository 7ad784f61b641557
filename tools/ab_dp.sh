#!/bin/bash
# Same-box A/B of the data-parallel transports (run under gpurun --gpus N): 1-GPU line, then peer / nccl alternating.
N=${1:-2}; STEPS=${2:-30}; MODE=${3:-pretrain}
mkdir -p gpurun_out
run() {  # tag transport gpus
  if [ "$3" = 1 ]; then
    python bench.py --steps $STEPS --warmup 3 --mode $MODE --no-cpu-baseline --no-gpu-baseline > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  else
    MH_DP_TRANSPORT=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $3 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $3 --steps $STEPS --warmup 3 --mode $MODE > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  fi
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{tag}.json").read().strip().splitlines()[-1])
    s = d["ms_per_step_series"]
    print(f"{tag:12s} n={d['n_gpus']} {d['value']/1e6:7.3f} M frames/s  {d['ms_per_step']:7.3f} ms  e2e {d['e2e']['ms_per_step']:7.3f} ms  first5 {sum(s[:5])/5:6.2f} last5 {sum(s[-5:])/5:6.2f}  clk {d['clocks']['sm_mhz']}", flush=True)
except Exception as e:
    print(tag, "FAILED", e)
    print(open(f"gpurun_out/ab_{tag}.err").read()[-1500:])
PY
}
run one_a x 1
run peer_a peer $N
run peersm_a peer-sm $N
run nccl_a nccl $N
run peer_b peer $N
run none_a none $N
run one_b x 1
