"""The C3 tree snapshot of bench.py, grown by the ORACLE on the CPU ('65 536 samples/round against a
4096-node oracle-grown snapshot'): car at rest at the origin, goal (100, 0, 0, 0), vmax 5, the 1000 boxes of config C3,
snapshot rounds of 2048 samples from srand(1) until the tree has 4096 nodes.  Writes tests/golden/c3_snapshot.npz
(4096 x 20 doubles, the node-record format of tests/cpulib.py).  Both bench arms load this file, so neither needs the
other's code to build the workload.  Takes a few minutes on one core."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cpulib import CpuPlanner, scene_c3_boxes  # noqa: E402

CAR, GOAL, VMAX, NODES, K = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0), (100.0, 0.0, 0.0, 0.0), 5.0, 4096, 2048

kind = sys.argv[1] if len(sys.argv) > 1 else "oracle"
orc = CpuPlanner(kind)
orc.set_obstacles(scene_c3_boxes())
orc.srand(1)
orc.tree_init(CAR, GOAL, VMAX)
t0 = time.time()
rounds = 0
while orc.tree_size() < NODES:
    s, h, _ = orc.draw_samples(K)
    orc.expand_round(s, h)
    rounds += 1
    print(f"round {rounds}: {orc.tree_size()} nodes, {time.time() - t0:.0f} s", flush=True)
tree = orc.tree_export()[:NODES]
out = os.path.join(ROOT, "tests", "golden", "c3_snapshot.npz")
np.savez_compressed(out, tree=tree, car=np.array(CAR), goal=np.array(GOAL), vmax=VMAX, samples_per_round=K, rounds=rounds,
                    grown_by=kind)
print("wrote", out, tree.shape)
