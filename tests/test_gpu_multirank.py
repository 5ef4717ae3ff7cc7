"""Two ranks on two GPUs (NCCL): sharded rounds + record all-gather give the single-GPU tree bit for bit.
Skipped on boxes with fewer than 2 GPUs (the world-size-2 host logic is covered on CPU by test_multirank_cpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_tree_equals_single_gpu_tree():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "scripts", "multirank_check.py")], capture_output=True, text=True, timeout=600)
    print(out.stdout[-2000:], out.stderr[-2000:])
    assert "MULTIRANK_OK" in out.stdout
