"""C5 debugging, CPU side: reference counters of query Q of the G5 loop after m = 0..100 expandTree calls (replays the
loop from the start for every m) -> variants/c5_ref_qQ.npy.  Usage: python scripts/c5_ref_side.py Q"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cpulib import CpuPlanner
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
g = np.load(os.path.join(ROOT, "tests/golden/g5_replan.npz"))
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 41
out = []
for m in range(0, 101):
    ref = CpuPlanner("ref_defined"); ref.srand(1); ref.commit_reset()
    for q in range(Q):
        ref.set_obstacles(g["obstacles"][q]); ref.query_commit(g["world"][q], g["goal"][q], 5.0, 100)
    ref.set_obstacles(g["obstacles"][Q])
    carried, tree, nbest, steps, cost = ref.query_commit(g["world"][Q], g["goal"][Q], 5.0, m)
    c = ref.counters()
    out.append([m, carried, tree, steps, c["fail_collision"], c["fail_acclimit"], c["fail_iterlimit"], c["sim_count"]])
out = np.array(out)
np.save(os.path.join(ROOT, "variants", f"c5_ref_q{Q}.npy"), out)
print(out[[1, 2, 50, 100]])
