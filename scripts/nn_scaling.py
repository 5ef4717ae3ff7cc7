"""Candidate search time against tree size, with and without the spatial sort (diagnostic; CLRRT_NN_SORT_MIN)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import clrrt_b200 as clrrt
import bench
K = 4096
pl = clrrt.Planner(device=0, tree_capacity=(1 << 19) + 2 * K, max_round=K)
pl.set_query(bench.C1_CAR, bench.C1_GOAL, 5.0)
pl.set_obstacles(bench.scene_c1_boxes())
pl.tree_reset(clrrt.root_node(bench.C1_CAR))
clrrt.draw_samples(bench.C1_GOAL, 1, seed=1)
marks = [4096, 16384, 65536, 131072, 262144]
while marks and pl.tree_size() < (1 << 19):
    s, h = clrrt.draw_samples(bench.C1_GOAL, K)
    st = pl.expand_round(s, h)
    if pl.tree_size() >= marks[0]:
        best = min(pl.expand_round(s, h).ms_nearest for _ in range(3))
        print(f"tree {pl.tree_size()} nodes: nearest {best:.3f} ms for {K} samples (sort threshold {os.environ.get('CLRRT_NN_SORT_MIN', 'default')})")
        marks.pop(0)
