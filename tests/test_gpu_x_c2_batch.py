"""BASELINE config C2 as a parity case: CellViT-256, batch 8 x 1024^2 -- the largest row counts any kernel of the engine sees
(32,776 token rows, 8.4 M pixel rows) -- against the fp32 oracle evaluated on the same GPU. (Added after the round's last GPU
run; the file sorts after the other GPU test files on purpose.)"""
import pytest
import torch

from test_gpu_forward import test_forward_full_size_1024_matches_oracle_on_device as _full_size_case

pytestmark = pytest.mark.gpu


def test_forward_cellvit256_batch8_1024_matches_oracle_on_device():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    _full_size_case("ViT256", 8)
