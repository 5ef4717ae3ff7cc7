"""Exploratory parity report on a GPU box (not a test): prints mismatch statistics GPU vs golden/oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import clrrt_b200 as clrrt
from cpulib import CpuPlanner, scene_c1_boxes, scene_c3_boxes, check_candidate_lists

G = os.path.join(ROOT, "tests", "golden")
DISC = [7, 12, 13, 14, 15, 16, 20, 21, 22]
CONT = [0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 17, 18, 19]

def report(name, got, want):
    d = (got[:, DISC] != want[:, DISC])
    bad = np.where(d.any(1))[0]
    err = np.abs(got[:, CONT] - want[:, CONT]) / np.maximum(1.0, np.abs(want[:, CONT]))
    err[bad] = 0
    print(f"{name}: n={len(got)} discrete-mismatch rows={len(bad)} max rel err (matching rows)={np.nanmax(err):.3e}")
    for i in bad[:5]:
        print("   row", i, "cols", [DISC[c] for c in np.where(d[i])[0]], "got", got[i, DISC], "want", want[i, DISC])

t0 = time.time()
g = np.load(os.path.join(G, "g1_rollouts.npz"))
pl = clrrt.Planner(device=0, tree_capacity=1 << 16, max_round=1 << 13)
pl.set_query(g["car"], g["goal"], 5.0)
pl.tree_reset_records(g["tree"])
for name, obs in (("obs", g["obstacles"]), ("live", np.zeros((0, 7)))):
    pl.set_obstacles(obs)
    report(f"G1 {name}", clrrt.rollouts_as_table(pl.propagate_batch(g["parent"], g["samples"])), g[f"out_{name}"])
    report(f"G1 root {name}", clrrt.rollouts_as_table(pl.propagate_batch(g["parent_root"], g["samples_root"])), g[f"out_root_{name}"])
    n = len(g["parent_gb"])
    report(f"G1 gb {name}", clrrt.rollouts_as_table(pl.propagate_batch(g["parent_gb"], np.zeros((n, 2)), np.ones(n, np.uint8))), g[f"out_gb_{name}"])
print("t", time.time() - t0)
g4 = np.load(os.path.join(G, "g4_dense.npz"))
pl.set_query(g4["car"], g4["goal"], 5.0)
pl.set_obstacles(g4["obstacles"])
pl.tree_reset_records(g4["tree"])
report("G4 dense", clrrt.rollouts_as_table(pl.propagate_batch(g4["parent"], g4["samples"])), g4["out"])
# trajectories
out, traj = pl.propagate_batch(g4["parent"][:4], g4["samples"][:4], traj_stride=512)
print("traj rows", out["n_steps"], traj[0, :3, :3])

g2 = np.load(os.path.join(G, "g2_nearest.npz"))
pl.set_query((0, 0, 0, 0, 3, 0), (50, 0, 0, 0), 5.0)
pl.set_obstacles(scene_c1_boxes())
orc = CpuPlanner("oracle"); orc.set_obstacles(scene_c1_boxes()); orc.tree_init((0, 0, 0, 0, 3, 0), (50, 0, 0, 0), 5.0)
for N in (1, 64, 250, 1500):
    pl.tree_reset_records(g2["tree"][:N]); orc.tree_import(g2["tree"][:N])
    cand, key, cnt = pl.nearest_batch(g2["samples"], g2["heuristic"])
    ocand, okey, ocnt = orc.nearest_batch(g2["samples"], g2["heuristic"])
    print(f"G2 N={N}: count equal {np.array_equal(cnt, ocnt)}; keys bit-equal {np.mean(key == okey):.5f}; cand equal {np.mean(cand == ocand):.5f};"
          f" max key ulp diff {np.max(np.abs(key.view(np.int32).astype(np.int64) - okey.view(np.int32).astype(np.int64)))}")
    try:
        check_candidate_lists(orc, g2["samples"], g2["heuristic"], cand, key, cnt); print("   spec check passed")
    except AssertionError as e:
        print("   spec check failed:", str(e)[:300])

g3 = np.load(os.path.join(G, "g3_replay.npz"))
for name, obs in (("live", np.zeros((0, 7))), ("obs", scene_c1_boxes())):
    pl.set_query((0, 0, 0, 0, 0, 0), (50, 0, 0, 0), 5.0)
    pl.set_obstacles(obs)
    pl.tree_reset(clrrt.root_node((0, 0, 0, 0, 0, 0)))
    t0 = time.time()
    for j in range(200):
        pl.expand_round(g3["samples"][j:j + 1], g3["heuristic"][j:j + 1])
    dt = time.time() - t0
    t = pl.tree_download_records(); w = g3[f"tree_{name}"]
    m = min(len(t), len(w))
    eq = [(np.array_equal(t[i, [7, 17, 18, 19]], w[i, [7, 17, 18, 19]]) and np.allclose(t[i], w[i], rtol=1e-6, atol=1e-9)) for i in range(m)]
    first_bad = eq.index(False) if False in eq else -1
    print(f"G3 {name}: gpu nodes {len(t)} golden {len(w)} first mismatch {first_bad} counters {pl.counters()} golden {g3[f'counters_{name}']} ({dt:.2f}s for 200 K=1 rounds) best {pl.best_path()} golden {g3[f'best_{name}']}")
    if first_bad >= 0: print(t[first_bad]); print(w[first_bad])
# a K=256 round vs the oracle's snapshot round
orc = CpuPlanner("oracle"); orc.set_obstacles(scene_c1_boxes()); orc.srand(5); orc.tree_init((0, 0, 0, 0, 2, 0)); orc.expand(50)
pl.set_query((0, 0, 0, 0, 2, 0), (50, 0, 0, 0), 5.0); pl.set_obstacles(scene_c1_boxes()); pl.tree_reset_records(orc.tree_export())
for r in range(3):
    s, h, _ = orc.draw_samples(256)
    orc.expand_round(s, h); st = pl.expand_round(s, h)
    a, b = pl.tree_download_records(), orc.tree_export()
    same_disc = len(a) == len(b) and np.array_equal(a[:, [7, 17, 18, 19]], b[:, [7, 17, 18, 19]])
    print(f"round {r}: gpu {len(a)} oracle {len(b)} discrete equal {same_disc} allclose {len(a)==len(b) and np.allclose(a, b, rtol=1e-6, atol=1e-9)} stats nodes+{st.nodes_added} steps {st.sim_steps} ms {st.ms_nearest:.3f}/{st.ms_rollout:.3f}/{st.ms_prepare:.3f}/{st.ms_append:.3f}")
# C3-size timing
boxes = scene_c3_boxes()
orc = CpuPlanner("oracle"); orc.set_obstacles(boxes); orc.srand(1); orc.tree_init((0, 0, 0, 0, 3, 0), (100, 0, 0, 0), 5.0); orc.expand(300)
pl2 = clrrt.Planner(device=0, tree_capacity=1 << 20, max_round=1 << 16)
pl2.set_query((0, 0, 0, 0, 3, 0), (100, 0, 0, 0), 5.0); pl2.set_obstacles(boxes); pl2.tree_reset_records(orc.tree_export())
n0 = pl2.tree_size()
s, h, _ = orc.draw_samples(65536)
for r in range(3):
    st = pl2.expand_round(s, h); pl2.tree_truncate(n0)
    print(f"C3 round: K=65536 tree {n0} nodes+{st.nodes_added} rollouts {st.rollouts} steps {st.sim_steps} ms nearest {st.ms_nearest:.3f} rollout {st.ms_rollout:.3f} goal {st.ms_prepare:.3f} append {st.ms_append:.3f} -> {st.sim_steps/ (st.ms_rollout+st.ms_prepare) * 1e3:.3e} steps/s")
pl2.set_obstacles(scene_c1_boxes())
for r in range(2):
    st = pl2.expand_round(s, h); pl2.tree_truncate(n0)
    print(f"C2-like round (10 boxes): rollouts {st.rollouts} steps {st.sim_steps} ms nearest {st.ms_nearest:.3f} rollout {st.ms_rollout:.3f} goal {st.ms_prepare:.3f} -> {st.sim_steps/ (st.ms_rollout+st.ms_prepare) * 1e3:.3e} steps/s")
