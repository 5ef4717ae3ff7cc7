"""What the candidate search does per sample on a large C1 tree, from the counters of a -DCLRRT_NN_STATS build
(CLRRT_LIB=variants/nnstats.so; diagnostic)."""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import clrrt_b200 as clrrt, bench
K = 16384
ROUNDS = int(os.environ.get("ROUNDS", "24"))
pl = clrrt.Planner(device=0, tree_capacity=(1 << 20) + 2 * K, max_round=K)
pl.set_query(bench.C1_CAR, bench.C1_GOAL, bench.VMAX); pl.set_obstacles(bench.scene_c1_boxes())
pl.tree_reset(clrrt.root_node(bench.C1_CAR))
clrrt.draw_samples(bench.C1_GOAL, 1, seed=1)
for r in range(ROUNDS):
    s, h = clrrt.draw_samples(bench.C1_GOAL, K); pl.expand_round(s, h)
s, h = clrrt.draw_samples(bench.C1_GOAL, K)
out = (C.c_ulonglong * 8)()
names = ["tile steps/block", "tiles loaded/block", "tiles searched/sample", "stage-1 survivors/sample", "feasible/sample", "inserted/sample"]
for name, hh in (("explore", np.zeros_like(h)), ("optimise", np.ones_like(h))):
    pl.lib.clrrt_debug_nn_stats(out, 1)
    pl.nearest_batch(s, hh)
    pl.lib.clrrt_debug_nn_stats(out, 1)
    v = list(out); nb = K / 8
    print(f"{name}: tree {pl.tree_size()} ({(pl.tree_size() + 255) // 256} tiles): " + ", ".join(
        f"{n} {v[i] / (nb if i < 2 else K):.1f}" for i, n in enumerate(names)))
