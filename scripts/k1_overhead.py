"""Where the 200 ms of the sequential C1 query go: wall clock, time inside the library, device phases, equal-key routes (diagnostic)."""
import os, sys, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import clrrt_b200 as clrrt, bench
from bench import C1_CAR, C1_GOAL, VMAX, scene_c1_boxes
for chunk in (32, 128):
    pl = clrrt.Planner(device=0, tree_capacity=1 << 16, max_round=256)
    pl.set_query(C1_CAR, C1_GOAL, VMAX); pl.set_obstacles(scene_c1_boxes()); pl.tree_reset(clrrt.root_node(C1_CAR))
    s, h = clrrt.draw_samples(C1_GOAL, 64, seed=1); pl.expand_sequential(s, h); pl.tree_reset(clrrt.root_node(C1_CAR))
    clrrt.draw_samples(C1_GOAL, 1, seed=1)
    it = windows = 0; dev = 0.0; tot = 0.0; tdraw = 0.0
    t0 = time.perf_counter()
    while (time.perf_counter() - t0) * 1e3 < 200.0:
        a = time.perf_counter(); s, h = clrrt.draw_samples(C1_GOAL, chunk); tdraw += time.perf_counter() - a
        st = pl.expand_sequential(s, h)
        it += chunk; windows += st.windows; fb = globals().get("fb", 0) + st.exact_fallbacks; globals()["fb"] = fb; sm = globals().get("sm", 0) + st.tie_checks_same; globals()["sm"] = sm; dev += st.ms_search + st.ms_prepare + st.ms_rollout + st.ms_commit; tot += st.ms_total
    wall = (time.perf_counter() - t0) * 1e3
    print(f"chunk {chunk}: fallbacks {fb} same-outcome {sm} iterations {it} windows {windows} wall {wall:.1f} ms; inside the library {tot:.1f} ms; device phases {dev:.1f} ms; drawing {tdraw*1e3:.1f} ms; per window: library {tot/windows*1e3:.0f} us device {dev/windows*1e3:.0f} us")
    pl.close()
