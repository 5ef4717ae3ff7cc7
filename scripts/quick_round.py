"""One-line timing of the C3 round (and K=4096) for the library selected by CLRRT_LIB (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import clrrt_b200 as clrrt
import bench
prm = clrrt.default_params(); prm.fp32 = 1 if os.environ.get('CLRRT_FP32') else 0
pl = clrrt.Planner(params=prm, device=0, tree_capacity=bench.TREE_SNAPSHOT + 2 * bench.K_ROUND + 1024, max_round=bench.K_ROUND)
boxes, smp, heu = bench.build_workload(pl, clrrt, 0, 1)
if os.environ.get('REFILL_MIN'): pl.set_tuning(refill_min=int(os.environ['REFILL_MIN']))
n0 = pl.tree_size()
tag = os.environ.get("CLRRT_LIB", "default")
for K in (bench.K_ROUND, 4096):
    best = None
    for r in range(4):
        st = pl.expand_round(smp[:K], heu[:K]); pl.tree_truncate(n0)
        t = (st.ms_nearest, st.ms_rollout, st.ms_prepare, st.ms_append)
        if best is None or sum(t) < sum(best): best = t
    print(f"{tag}: K={K} rollouts {st.rollouts} steps {st.sim_steps} added {st.nodes_added} ms nearest {best[0]:.2f} rollout {best[1]:.2f} goal {best[2]:.2f} append {best[3]:.2f} total {sum(best):.2f}")
