"""fp32 mode vs the fp64 golden rollouts: verdict agreement and state error (diagnostic, GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import clrrt_b200 as clrrt
import bench
G = os.path.join(ROOT, "tests", "golden")
g = np.load(os.path.join(G, "g1_rollouts.npz"))
p = clrrt.default_params(); p.fp32 = 1
pl = clrrt.Planner(params=p, device=0, tree_capacity=1 << 16, max_round=1 << 16)
for name, obs in (("obs", g["obstacles"]), ("live", np.zeros((0, 7)))):
    pl.set_query(g["car"], g["goal"], 5.0); pl.tree_reset_records(g["tree"]); pl.set_obstacles(obs)
    got = clrrt.rollouts_as_table(pl.propagate_batch(g["parent"], g["samples"])); want = g[f"out_{name}"]
    same = (got[:, 15] == want[:, 15]) & (got[:, 12] == want[:, 12]) & (got[:, 13] == want[:, 13])
    acc = same & ((want[:, 12] + want[:, 13]) > 0)
    dpos = np.hypot(got[:, 0] - want[:, 0], got[:, 1] - want[:, 1])
    print(f"fp32 G1 {name}: verdict agreement {same.mean():.4f} ({(~same).sum()} of {len(same)} differ); accepted&agreeing {acc.sum()}: "
          f"pos err max {dpos[acc].max():.2e} p99 {np.percentile(dpos[acc], 99):.2e} m, heading max {np.abs(got[acc,2]-want[acc,2]).max():.2e}, "
          f"steps diff max {np.abs(got[acc,14]-want[acc,14]).max():.0f}, costE rel {np.abs(got[acc,10]/want[acc,10]-1).max():.2e} costS rel {np.abs(got[acc,11]/want[acc,11]-1).max():.2e}")
    bad = np.where(~same)[0][:6]
    for i in bad: print("   differ:", i, "fp32 fail", got[i, 15], "steps", got[i, 14], "| fp64 fail", want[i, 15], "steps", want[i, 14])
# throughput
pl2 = clrrt.Planner(params=p, device=0, tree_capacity=bench.TREE_SNAPSHOT + 2 * bench.K_ROUND + 1024, max_round=bench.K_ROUND)
boxes, smp, heu = bench.build_workload(pl2, clrrt, 0, 1)
n0 = pl2.tree_size()
for r in range(3):
    st = pl2.expand_round(smp, heu); pl2.tree_truncate(n0)
print(f"fp32 C3 round: rollouts {st.rollouts} steps {st.sim_steps} nodes+{st.nodes_added} ms nearest {st.ms_nearest:.2f} rollout {st.ms_rollout:.2f} goal {st.ms_prepare:.2f} -> {st.sim_steps/(st.ms_nearest+st.ms_rollout+st.ms_prepare+st.ms_append)*1e3:.3e} steps/s")
