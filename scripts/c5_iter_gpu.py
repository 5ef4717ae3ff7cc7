"""C5 debugging, GPU side: runs the rollouts recorded by scripts/c5_iter_cpu.py on the device and reports the first sim
step at which a trajectory leaves the oracle's.  Usage (under gpurun): python scripts/c5_iter_gpu.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import clrrt_b200 as clrrt
d = np.load(os.path.join(ROOT, "variants", "c5_iter.npz"))
pl = clrrt.Planner(device=0, tree_capacity=4096, max_round=64)
pl.set_query(d["car"], d["goal"], 5.0)
pl.set_obstacles(d["obstacles"])
pl.tree_reset_records(d["tree"])
par, gb = d["par"], d["gb"]
smp = np.repeat(d["sample"], len(par), 0)
out, traj = pl.propagate_batch(par, smp, gb, traj_stride=512)
for i in range(len(par)):
    want = d[f"traj{i}"]
    n = int(out["n_steps"][i])
    got = traj[i, :n + 1]
    print(f"rollout {i} parent {par[i]} gb {gb[i]}: ours rows {n+1} fail {out['fail'][i]} end {out['end_reached'][i]} goal {out['goal_reached'][i]}; oracle rows {len(want)}")
    m = min(len(want), len(got))
    err = np.abs(got[:m, :7] - want[:m, :7])
    bad = np.where(err.max(axis=1) > 1e-9)[0]
    if len(bad):
        k = bad[0]
        print("   first row differing by > 1e-9:", k, "of", m, "\n   ours  ", got[k], "\n   oracle", want[k])
        for kk in range(max(0, k - 2), min(m, k + 3)):
            print("   row", kk, "max err", err[kk].max(), "idwp ours/oracle", got[kk, 7], want[kk, 7], "dcmd", got[kk, 9], want[kk, 9])
    if len(want) != len(got):
        print("   last rows ours", got[-1][:8], "\n   last rows oracle", want[-1][:8])
