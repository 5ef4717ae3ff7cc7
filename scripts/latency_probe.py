"""Per-step latency of a lone rollout (one lane of one warp on an otherwise idle GPU) — diagnostic."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import clrrt_b200 as clrrt
import bench
pl = clrrt.Planner(device=0, tree_capacity=1 << 14, max_round=1 << 12)
car, goal = bench.CAR, bench.GOAL
pl.set_query(car, goal, 5.0)
pl.tree_reset(clrrt.root_node(car))
def probe(tag, par, smp, gb, reps=20):
    out = pl.propagate_batch(par, smp, gb)
    t0 = time.perf_counter()
    for _ in range(reps): out = pl.propagate_batch(par, smp, gb)
    dt = (time.perf_counter() - t0) / reps
    steps = out["n_steps"]
    print(f"{tag}: M={len(par)} max steps {steps.max()} total {steps.sum()} call {dt*1e6:.0f} us -> {dt*1e6/max(1,steps.max()):.2f} us per step of the longest rollout")
for nobs, boxes in (("no obstacles", np.zeros((0, 7))), ("1000 obstacles", bench.scene_c3_boxes())):
    pl.set_obstacles(boxes)
    # a long straight rollout down the corridor from the root
    probe(f"{nobs}, 1 rollout", [0], [[60.0, 0.0]], [0])
    probe(f"{nobs}, 32 rollouts (1 warp)", [0] * 32, [[60.0, 0.1 * i - 1.5] for i in range(32)], [0] * 32)
    probe(f"{nobs}, 4096 rollouts", [0] * 4096, [[60.0, 0.0007 * i - 1.5] for i in range(4096)], [0] * 4096)
    probe(f"{nobs}, 1 goal-biased rollout", [0], [[0.0, 0.0]], [1])
probe("empty call overhead (1 one-step rollout)", [0], [[0.5, 0.0]], [0])
