"""Where do goal-biased rollouts start to differ from the oracle? (diagnostic, GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import clrrt_b200 as clrrt
from cpulib import CpuPlanner
np.set_printoptions(linewidth=220, precision=9)
G = os.path.join(ROOT, "tests", "golden")
g = np.load(os.path.join(G, "g1_rollouts.npz"))
pl = clrrt.Planner(device=0, tree_capacity=1 << 12, max_round=1 << 10)
pl.set_query(g["car"], g["goal"], 5.0); pl.tree_reset_records(g["tree"]); pl.set_obstacles(np.zeros((0, 7)))
orc = CpuPlanner("oracle"); orc.set_obstacles(np.zeros((0, 7))); orc.tree_init(g["car"], g["goal"], 5.0); orc.tree_import(g["tree"])
n = len(g["parent_gb"])
out, traj = pl.propagate_batch(g["parent_gb"], np.zeros((n, 2)), np.ones(n, np.uint8), traj_stride=512)
want = g["out_gb_live"]
got = clrrt.rollouts_as_table(out)
for i in range(n):
    ot, _ = orc.rollout_traj(int(g["parent_gb"][i]), [0, 0], 1)
    m = len(ot)
    gt = traj[i, :m]
    rel = np.abs(gt - ot) / np.maximum(1, np.abs(ot))
    bad = np.where(rel.max(1) > 1e-9)[0]
    colerr = np.abs(got[i] - want[i]) / np.maximum(1, np.abs(want[i]))
    print(f"gb {i}: steps {m-1} Nref {int(want[i,16])} first step >1e-9: {bad[0] if len(bad) else -1} idwp there {int(ot[bad[0],7]) if len(bad) else -1} "
          f"max final relerr {colerr.max():.2e} at col {colerr.argmax()}")
    if len(bad) and i < 3:
        b = bad[0]
        print("   gpu", gt[b - 1:b + 2, [0, 1, 2, 3, 7, 9]].tolist()); print("   orc", ot[b - 1:b + 2, [0, 1, 2, 3, 7, 9]].tolist())
