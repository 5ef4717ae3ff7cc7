"""Where the 200 ms query (config C1) spends its time: per round, tree size and the library's phase timings (diagnostic)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import clrrt_b200 as clrrt
import bench
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
pl = clrrt.Planner(device=0, tree_capacity=(1 << 20) + 2 * K, max_round=K)
pl.set_query(bench.C1_CAR, bench.C1_GOAL, bench.VMAX); pl.set_obstacles(bench.scene_c1_boxes()); pl.tree_reset(clrrt.root_node(bench.C1_CAR))
s, h = clrrt.draw_samples(bench.C1_GOAL, K, seed=1); pl.expand_round(s, h); pl.tree_reset(clrrt.root_node(bench.C1_CAR))
ROUNDS = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # > 0: a fixed number of rounds instead of the 200 ms budget
t0 = time.perf_counter(); r = 0
while (r < ROUNDS) if ROUNDS else ((time.perf_counter() - t0) * 1e3 < 200.0):
    ta = time.perf_counter(); s, h = clrrt.draw_samples(bench.C1_GOAL, K); tb = time.perf_counter()
    st = pl.expand_round(s, h); tc = time.perf_counter()
    print(f"round {r:2d}: tree {st.tree_size:7d} (+{st.nodes_added}) draw {1e3*(tb-ta):.2f} ms call {1e3*(tc-tb):.2f} ms = nearest {st.ms_nearest:.2f} prepare {st.ms_prepare:.2f} rollout {st.ms_rollout:.2f} append {st.ms_append:.2f}; steps {st.sim_steps}")
    r += 1
