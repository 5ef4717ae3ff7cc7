"""Config C4 at its full size on one GPU: a round of 2^20 samples (10 boxes, 4096-node snapshot) expanded at once, and the
same round expanded as 8 contiguous shards with append deferred — what 8 ranks do — whose record chunks are appended in
shard order: the trees must be byte-identical (the multi-GPU exchange moves exactly these chunks, tests/test_gpu_multirank.py
and scripts/multirank_check.py cover the NCCL leg).  A 512-sample subset is checked against the oracle as in the C3 test."""
import numpy as np
import pytest

import clrrt_b200 as clrrt
from clrrt_b200.exchange import shard_range
from cpulib import CpuPlanner, scene_c1_boxes

pytestmark = pytest.mark.gpu


def test_c4_round_full_size_sharded_equals_whole():
    import torch
    import bench
    K, WORLD = 1 << 20, 8
    car, goal = (0.0, 0.0, 0.0, 0.0, 3.0, 0.0), (50.0, 0.0, 0.0, 0.0)
    boxes = scene_c1_boxes()
    pl = clrrt.Planner(device=0, tree_capacity=4096 + 2 * K + 1024, max_round=K)
    try:
        pl.set_query(car, goal, 5.0)
        pl.set_obstacles(boxes)
        pl.tree_reset(clrrt.root_node(car))
        s0, h0 = clrrt.draw_samples(goal, 8192 * 8, seed=1)
        i = 0
        while pl.tree_size() < 4096:
            pl.expand_round(s0[i * 8192:(i + 1) * 8192], h0[i * 8192:(i + 1) * 8192]); i += 1
        pl.tree_truncate(4096)
        n0 = 4096
        snapshot = pl.tree_download_records()
        smp, heu = clrrt.draw_samples(goal, K, seed=2)
        st = pl.expand_round(smp, heu)
        whole = pl.tree_download()[n0:].tobytes()
        n_whole = st.nodes_added
        assert n_whole > 100000
        whole_rec = pl.tree_download_records()[n0:]
        # the same round as 8 shards, append deferred, chunks appended in shard order
        pl.tree_truncate(n0)
        pl.set_defer_append(True)
        chunks, counts = [], []
        for r in range(WORLD):
            lo, hi = shard_range(K, r, WORLD)
            pl.expand_round(smp[lo:hi], heu[lo:hi])
            ptr, n = pl.round_records()
            chunks.append(bench._as_cuda_tensor(ptr, n * clrrt.RECORD_BYTES, 0).clone() if n else torch.empty(0, dtype=torch.uint8, device="cuda"))
            counts.append(n)
        stride = max(counts)
        buf = torch.zeros(WORLD * stride * clrrt.RECORD_BYTES, dtype=torch.uint8, device="cuda")
        for r, c in enumerate(chunks):
            buf[r * stride * clrrt.RECORD_BYTES: r * stride * clrrt.RECORD_BYTES + c.numel()] = c
        torch.cuda.synchronize()
        pl.append_records(buf.data_ptr(), np.array(counts, np.int32), stride)
        pl.set_defer_append(False)
        assert pl.tree_size() == n0 + n_whole == n0 + sum(counts)
        assert pl.tree_download()[n0:].tobytes() == whole, "sharded round differs from the whole round"
        # a subset against the oracle
        sel = np.sort(np.random.default_rng(9).choice(K, 512, replace=False))
        orc = CpuPlanner("oracle")
        orc.set_obstacles(boxes)
        orc.tree_init(car, goal, 5.0)
        orc.tree_import(snapshot)
        orc.expand_round(smp[sel], heu[sel])
        want = orc.tree_export()[n0:]
        key = {tuple(r[:7]) + (r[12], r[13]): i for i, r in enumerate(whole_rec)}
        last = -1
        for r in want:
            i = key.get(tuple(r[:7]) + (r[12], r[13]))
            assert i is not None and i > last
            last = i
            assert np.array_equal(whole_rec[i][7:17], r[7:17]) and whole_rec[i][18] == r[18] and whole_rec[i][19] == r[19]
        print(f"C4 full size: {n_whole} nodes from {K} samples; 8 shards + ordered append byte-identical to the whole round; "
              f"{len(want)} nodes of a 512-sample subset bit-equal to the oracle's")
    finally:
        pl.close()
