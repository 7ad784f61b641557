"""torchrun smoke of train.py in data-parallel mode (default: the step incl. NCCL captured into a CUDA graph):
head-pruning rebuilds the step (new shapes, new capture) twice in six optimizer steps.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_train_smoke.py [mode]
"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from test_gpu_entrypoints import _write_cfgs
import train

mode = sys.argv[1] if len(sys.argv) > 1 else "head-pruning"
rank = int(os.environ.get("RANK", "0"))
tmp = os.path.join(tempfile.gettempdir(), f"dp_smoke_{mode}_{rank}")
os.makedirs(tmp, exist_ok=True)
mp, rp = _write_cfgs(tmp, mode)
exp = os.path.join(tmp, "exp")
train.main(["-m", mode, "-g", mp, "-c", rp, "-n", exp, "-f", "20", "--synthetic", "--multi_gpu"])
torch.cuda.synchronize()
if rank == 0:
    st = torch.load(os.path.join(exp, "last-step.ckpt"), map_location="cpu", weights_only=False)
    print("[dp_train_smoke]", mode, "steps", st["Step"], "q_proj", tuple(st["model"]["encoder.layers.0.self_attn.q_proj.weight"].shape),
          open(os.path.join(exp, "train_log.csv")).read().strip().splitlines()[-1])
