#!/bin/bash
# Is the N = 2 overhead the exchange or the box?  (a) two INDEPENDENT 1-GPU benches side by side, (b) 2 ranks with the
# gradient exchange disabled (loss all-reduce only), (c) 2 ranks with the peer transport.
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{tag}.json").read().strip().splitlines()[-1])
    s = d["ms_per_step_series"]
    print(f"{tag:12s} n={d['n_gpus']} {d['ms_per_step']:7.3f} ms  first5 {sum(s[:5])/5:6.2f} last5 {sum(s[-5:])/5:6.2f}  clk {d['clocks']['sm_mhz']}", flush=True)
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/ab_{tag}.err").read()[-800:])
PY
}
A="--steps 30 --warmup 3 --no-cpu-baseline --no-gpu-baseline"
python bench.py $A > gpurun_out/ab_solo.json 2> gpurun_out/ab_solo.err; show solo
CUDA_VISIBLE_DEVICES=0 python bench.py $A > gpurun_out/ab_side0.json 2> gpurun_out/ab_side0.err &
CUDA_VISIBLE_DEVICES=1 python bench.py $A > gpurun_out/ab_side1.json 2> gpurun_out/ab_side1.err &
wait; show side0; show side1
CUDA_VISIBLE_DEVICES=1 python bench.py $A > gpurun_out/ab_solo1.json 2> gpurun_out/ab_solo1.err; show solo1
for t in none peer; do
MH_DP_TRANSPORT=$t python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/ab_dp_$t.json 2> gpurun_out/ab_dp_$t.err; show dp_$t
done
