"""Tuning sweep of the C3 round on a GPU box: persistent-grid size and refill threshold (diagnostic)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import clrrt_b200 as clrrt
import bench

pl = clrrt.Planner(device=0, tree_capacity=bench.TREE_SNAPSHOT + 2 * bench.K_ROUND + 1024, max_round=bench.K_ROUND)
boxes, smp, heu = bench.build_workload(pl, clrrt, 0, 1)
n0 = pl.tree_size()
def run(tag, K=bench.K_ROUND, reps=2):
    for r in range(reps):
        st = pl.expand_round(smp[:K], heu[:K]); pl.tree_truncate(n0)
    print(f"{tag}: K={K} rollouts {st.rollouts} steps {st.sim_steps} ms nearest {st.ms_nearest:.2f} rollout {st.ms_rollout:.2f} goal {st.ms_prepare:.2f} append {st.ms_append:.2f} -> {st.sim_steps/(st.ms_rollout+st.ms_prepare)*1e3:.3e} steps/s (kernels), {st.sim_steps/(st.ms_rollout)*1e3:.3e} main only")
for bps in (0,):
    for rf in (1, 8, 16, 32):
        pl.set_tuning(refill_min=rf, blocks_per_sm=bps)
        run(f"blocks/SM={bps} refill_min={rf}")
pl.set_tuning(refill_min=8, blocks_per_sm=0)
for sl in (0.5, 1.0, 2.0, 4.0):
    pl.set_grid_cell(sl); pl.set_obstacles(boxes); run(f"grid cell {sl}")
pl.set_grid_cell(1.0); pl.set_obstacles(boxes)
pl.set_tuning(refill_min=1, blocks_per_sm=0)
for K in (4096, 16384, 65536):
    run("K sweep", K)
# chain statistics from the candidate lists and single rollouts of the first 8192 samples
cand, key, cnt = pl.nearest_batch(smp[:8192], heu[:8192])
par = np.concatenate([cand[j, :cnt[j]] for j in range(8192)]); sm = np.concatenate([np.repeat(smp[j:j+1], cnt[j], 0) for j in range(8192)])
out = pl.propagate_batch(par, sm)
steps = out["n_steps"]; ok = (out["end_reached"] + out["goal_reached"]) > 0
chain = []; i = 0
for j in range(8192):
    tot = 0
    for r in range(cnt[j]):
        tot += steps[i + r]
        if ok[i + r]: break
    chain.append(tot); i += cnt[j]
chain = np.array(chain)
print("rollout steps: mean %.1f p50 %d p90 %d p99 %d max %d; fail codes %s" % (steps.mean(), *np.percentile(steps, [50, 90, 99]), steps.max(), np.bincount(out["fail"], minlength=4)))
print("chain steps: mean %.1f p50 %d p90 %d p99 %d max %d; candidates/sample mean %.2f" % (chain.mean(), *np.percentile(chain, [50, 90, 99]), chain.max(), cnt.mean()))
