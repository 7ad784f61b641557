"""Real-NCCL data-parallel equivalence: needs >= 2 GPUs (``gpurun --gpus 2``); skipped on one."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport", ["peer", "peer-sm", "nccl"])
def test_data_parallel_gradients_equal_single_gpu_and_prune_decisions_agree(transport):
    """All bucket transports: the own peer-memory reduce-scatter / all-gather (csrc/peer.cu: copy-engine pulls, the
    default, and SM pulls) and NCCL all-reduce."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, MH_DP_TRANSPORT=transport))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(": OK") == 2
