"""C5 debugging, CPU side: reproduces expandTree call M+1 of query Q with the oracle on the reference's tree (candidates,
every rollout of the call step by step) -> variants/c5_iter.npz.  Usage: python scripts/c5_iter_cpu.py Q M"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cpulib import CpuPlanner
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
g = np.load(os.path.join(ROOT, "tests/golden/g5_replan.npz"))
Q, M = int(sys.argv[1]), int(sys.argv[2])
ref = CpuPlanner("ref_defined"); ref.srand(1); ref.commit_reset()
for q in range(Q):
    ref.set_obstacles(g["obstacles"][q]); ref.query_commit(g["world"][q], g["goal"][q], 5.0, 100)
ref.set_obstacles(g["obstacles"][Q])
carried, tree_n, nbest, steps, cost = ref.query_commit(g["world"][Q], g["goal"][Q], 5.0, M)
print("after", M, "iterations: tree", tree_n, "sim", steps)
tree = ref.tree_export()
smp, heu, _ = ref.draw_samples(1)
print("sample", smp, heu)
w = g["world"][Q]
car = np.array([0, 0, 0, w[3], w[4], w[5]])
goal = g["goal"][Q]
orc = CpuPlanner("oracle")
orc.set_obstacles(g["obstacles"][Q])
orc.tree_init(car, goal, 5.0)
orc.tree_import(tree)
cand, key, cnt = orc.nearest_batch(smp, heu)
print("candidates", cand[0, :cnt[0]], key[0, :cnt[0]])
c0 = orc.counters()
n_added = orc.expand_with(smp, heu)
c1 = orc.counters()
tree2 = orc.tree_export()
print("oracle expand_with: tree", len(tree), "->", len(tree2), "sim steps", c1["sim_count"] - c0["sim_count"], "(reference iteration 74: 206, ours 213)")
new = tree2[len(tree):]
print("new nodes: parent, goal, nref:", new[:, [17, 18, 19]])
# the rollouts of this iteration, step by step
rolls = []
orc2 = CpuPlanner("oracle"); orc2.set_obstacles(g["obstacles"][Q]); orc2.tree_init(car, goal, 5.0); orc2.tree_import(tree2)
for r in range(cnt[0]):
    tr, _ = orc2.rollout_traj(int(cand[0, r]), smp[0], 0)
    res = orc2.rollout_batch([cand[0, r]], smp)
    print("cand", r, "parent", cand[0, r], "rows", len(tr), "fail", res[0, 12:17])
    rolls.append((int(cand[0, r]), 0, tr))
    if res[0, 13] != 0 or res[0, 14] != 0 or res[0,12]==0: pass
for k in range(len(tree), len(tree2)):
    if tree2[k, 17] >= len(tree) or True:
        pass
# goal-biased child: parent = first new node
tr, _ = orc2.rollout_traj(len(tree), smp[0], 1)
print("goal-biased from node", len(tree), "rows", len(tr))
rolls.append((len(tree), 1, tr))
np.savez(os.path.join(ROOT, "variants", "c5_iter.npz"), tree=tree2, car=car, goal=goal, obstacles=g["obstacles"][Q], sample=smp, heur=heu,
         par=np.array([r[0] for r in rolls]), gb=np.array([r[1] for r in rolls]), **{f"traj{i}": r[2] for i, r in enumerate(rolls)})
