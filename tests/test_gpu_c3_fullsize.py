"""Config C3 at its full size (65 536 samples per round, 1000 oriented boxes, 4096-node snapshot — the workload bench.py
times), through properties that do not need the oracle to run the whole round:
  * the round's result does not depend on how it is scheduled or searched (launch tuning, occupancy, exhaustive against
    sorted candidate search): byte-identical trees;
  * samples of a round are independent given the snapshot, so the nodes of any subset of the samples — here 768 of them,
    which the oracle expands in seconds — must appear in the full round's result, bit for bit and in order."""
import numpy as np
import pytest

import clrrt_b200 as clrrt
from cpulib import CpuPlanner

pytestmark = pytest.mark.gpu


def test_c3_round_full_size():
    import bench
    K = bench.K_ROUND
    pl = clrrt.Planner(device=0, tree_capacity=bench.TREE_SNAPSHOT + 2 * K + 1024, max_round=K)
    try:
        boxes, smp, heu = bench.build_workload(pl, clrrt, 0, 1)
        n0 = pl.tree_size()
        snapshot = pl.tree_download_records()
        st = pl.expand_round(smp, heu)
        full = pl.tree_download_records()[n0:]
        assert st.nodes_added == len(full) > 20000 and st.sim_steps > 10_000_000
        # 1. scheduling / search variants
        for refill, blocks, mode in ((1, 1, 2), (32, 0, 1), (7, 2, 18)):
            pl.tree_truncate(n0)
            pl.set_tuning(refill_min=refill, blocks_per_sm=blocks)
            pl.set_nearest_mode(mode)
            st2 = pl.expand_round(smp, heu)
            again = pl.tree_download_records()[n0:]
            assert (st2.nodes_added, st2.sim_steps, st2.rollouts) == (st.nodes_added, st.sim_steps, st.rollouts)
            assert again.tobytes() == full.tobytes(), f"tuning {(refill, blocks, mode)} changed the round's result"
        pl.set_tuning(refill_min=4, blocks_per_sm=0)
        pl.set_nearest_mode(0)
        # 2. a subset of the samples against the oracle
        sel = np.sort(np.random.default_rng(5).choice(K, 768, replace=False))
        orc = CpuPlanner("oracle")
        orc.set_obstacles(boxes)
        orc.tree_init(bench.CAR, bench.GOAL, bench.VMAX)
        orc.tree_import(snapshot)
        orc.expand_round(smp[sel], heu[sel])
        want = orc.tree_export()[n0:]
        assert len(want) > 50
        # parents inside the snapshot are absolute ids; goal-biased children point at the node before them
        # (state alone is not a key: two samples on one ray from the same parent can end in the same state)
        key = {tuple(r[:7]) + (r[12], r[13]): i for i, r in enumerate(full)}
        last = -1
        for r in want:
            i = key.get(tuple(r[:7]) + (r[12], r[13]))
            assert i is not None, "a node of the subset round is missing from the full round"
            assert i > last, "order of the subset's nodes differs"
            last = i
            g = full[i]
            assert np.array_equal(g[7:17], r[7:17]) and g[18] == r[18] and g[19] == r[19]
            if r[17] < n0:
                assert g[17] == r[17]
            else:
                assert g[17] == n0 + i - 1  # the goal-biased child follows its parent
        print(f"C3 full size: {len(full)} nodes from {K} samples in {st.sim_steps} sim steps; 3 scheduling variants byte-identical; "
              f"{len(want)} nodes of a 768-sample subset bit-equal to the oracle's")
    finally:
        pl.close()
