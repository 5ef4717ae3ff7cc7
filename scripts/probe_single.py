"""One lone rollout (1 lane) through 1000 obstacles, for an ncu latency capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import clrrt_b200 as clrrt
import bench
pl = clrrt.Planner(device=0, tree_capacity=1 << 12, max_round=1 << 10)
pl.set_query(bench.CAR, bench.GOAL, 5.0)
pl.tree_reset(clrrt.root_node(bench.CAR))
pl.set_obstacles(bench.scene_c3_boxes())
for _ in range(3):
    out = pl.propagate_batch([0], [[60.0, 0.0]], [0])
print(out["n_steps"], out["fail"])
