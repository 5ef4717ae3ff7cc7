"""A lone rollout (one lane on an idle GPU) for ncu's source-level stall sampling: where does a step's latency go?
    ncu --set full --import-source on -k regex:rollout_kernel -s 2 -c 1 -o gpurun_out/lone -f python scripts/lone_profile.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import clrrt_b200 as clrrt
import bench
pl = clrrt.Planner(device=0, tree_capacity=1 << 14, max_round=1 << 12)
pl.set_query(bench.CAR, bench.GOAL, 5.0)
pl.set_obstacles(bench.scene_c3_boxes())
pl.tree_reset(clrrt.root_node(bench.CAR))
for _ in range(4):
    out = pl.propagate_batch([0], [[0.0, 0.0]], [1])
print(out["n_steps"], out["fail"])
