"""Two ranks on two GPUs through the library's own NCCL exchange (clrrt_comm_init, ncclAllGather inside
libclrrt_b200.so — no torch.distributed on the data plane): sharded rounds give the single-GPU tree bit for bit, equal
device digests on every rank, counters summed over ranks equal the single-GPU counters.  Skipped with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exchange_two_ranks_equal_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "multirank_lib_check.py"), "2"], capture_output=True,
                         text=True, timeout=900)
    print(out.stdout[-3000:], out.stderr[-3000:])
    assert "MULTIRANK_LIB_OK" in out.stdout


def test_digest_is_order_sensitive_and_reproducible():
    """One GPU: the digest of a tree equals the digest of the same tree uploaded again, and changes when two nodes swap."""
    import numpy as np
    import clrrt_b200 as clrrt
    from cpulib import scene_c1_boxes
    pl = clrrt.Planner(device=0, tree_capacity=4096, max_round=1024)
    pl.set_query((0, 0, 0, 0, 3, 0), (50, 0, 0, 0), 5.0)
    pl.set_obstacles(scene_c1_boxes())
    pl.tree_reset(clrrt.root_node((0, 0, 0, 0, 3, 0)))
    s, h = clrrt.draw_samples((50, 0, 0, 0), 1024, seed=3)
    pl.expand_round(s, h)
    d0 = pl.tree_digest()
    nodes = pl.tree_download()
    pl.tree_reset(nodes)
    assert pl.tree_digest() == d0
    assert pl.tree_digest(5, 100) != pl.tree_digest(6, 100)
    sw = nodes.copy()
    leaves = [i for i in range(1, len(sw)) if not (sw["parent"] == i).any()][:2]
    a, b = leaves
    sw[[a, b]] = sw[[b, a]]
    if sw["parent"][a] < a and sw["parent"][b] < b:
        pl.tree_reset(sw)
        assert pl.tree_digest() != d0
    pl.close()
