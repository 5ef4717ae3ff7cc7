"""Sweep of the launch knobs of the round kernel on the C3 workload (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import clrrt_b200 as clrrt
import bench
pl = clrrt.Planner(device=0, tree_capacity=bench.TREE_SNAPSHOT + 2 * bench.K_ROUND + 1024, max_round=bench.K_ROUND)
boxes, smp, heu = bench.build_workload(pl, clrrt, 0, 1)
n0 = pl.tree_size()
for bps in (0, 2, 1):
    for rf in (1, 2, 4, 8):
        pl.set_tuning(refill_min=rf, blocks_per_sm=bps)
        best = 1e9
        for r in range(3):
            st = pl.expand_round(smp, heu); pl.tree_truncate(n0)
            best = min(best, st.ms_rollout)
        print(f"blocks/SM={bps or 'max'} refill_min={rf}: rollout kernel {best:.2f} ms")
