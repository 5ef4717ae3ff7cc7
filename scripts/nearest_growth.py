"""Search / rollout time per snapshot round as the C1 tree grows (diagnostic: the candidate search on large trees)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import clrrt_b200 as clrrt
import bench
K = 16384
pl = clrrt.Planner(device=0, tree_capacity=(1 << 20) + 2 * K, max_round=K)
pl.set_query(bench.C1_CAR, bench.C1_GOAL, bench.VMAX)
pl.set_obstacles(bench.scene_c1_boxes())
pl.tree_reset(clrrt.root_node(bench.C1_CAR))
s, h = clrrt.draw_samples(bench.C1_GOAL, K, seed=1)
pl.expand_round(s, h)
pl.tree_reset(clrrt.root_node(bench.C1_CAR))
clrrt.draw_samples(bench.C1_GOAL, 1, seed=1)
for r in range(48):
    s, h = clrrt.draw_samples(bench.C1_GOAL, K)
    n = pl.tree_size()
    st = pl.expand_round(s, h)
    if r % 4 == 3 or r < 2:
        print(f"round {r}: tree {n} -> {pl.tree_size()} nearest {st.ms_nearest:.2f} rollout {st.ms_rollout:.2f} goal {st.ms_prepare:.2f} append {st.ms_append:.2f} ms")
# the two keys apart, on the final tree (wall clock around the host-buffer search call, so copies included)
import time, numpy as np
s, h = clrrt.draw_samples(bench.C1_GOAL, K)
for name, hh in (("explore", np.zeros_like(h)), ("optimise", np.ones_like(h)), ("mixed", h)):
    pl.nearest_batch(s, hh)
    t0 = time.perf_counter()
    for _ in range(3): pl.nearest_batch(s, hh)
    print(f"round search {name}: tree {pl.tree_size()} {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms")
