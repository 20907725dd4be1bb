"""N-GPU result equality over NCCL: tiles sharded over 2 ranks with one NCCL weight broadcast give the
same per-tile instance dicts and cell tokens as one GPU processing every tile. Needs two GPUs (skipped on a 1-GPU box;
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_over_nccl_equal_one_gpu():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_equal_worker.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "NCCL_EQUAL_OK" in r.stdout, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
