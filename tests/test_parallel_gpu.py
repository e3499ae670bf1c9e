"""Data parallel ON HARDWARE: 2 ranks, NCCL, the CUDA kernels.  Skipped on a box with one GPU (the CPU / gloo versions of the
same checks are in tests/test_parallel.py); run with `gpurun --gpus 2 -- python -m pytest tests/test_parallel_gpu.py -m gpu`."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("mode", ["hdiff", "ddp"])
def test_two_rank_nccl_gradients_match_single_gpu(mode, dtype):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_nccl_worker.py"), mode, dtype]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok (") == 2, r.stdout[-2000:]
