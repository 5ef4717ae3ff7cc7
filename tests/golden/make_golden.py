"""Generate tests/golden/*.npz from the REFERENCE's own sources (oracle/_ref, built by oracle/build_ref.sh
from /root/reference).  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Every array is produced by libclrrt_ref_defined.so (reference + the documented UB clamps) and
cross-checked here against the unmodified libclrrt_ref.so on all untainted rollouts (bit-for-bit); the mask of
rows where the unmodified build differs is stored as `unmod_differs` and must be a subset of `tainted`.
Golden sets follow SURVEY.md §8c: G0 known answers, G1 rollouts (C2), G2 candidate lists, G3 whole-query
replay at K=1, G4 dense-scene verdicts (C3), G5 receding-horizon loop with carried-over trees (C5), G6 curved-road cost (bend), G7 road-frame transforms and lane samples.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from cpulib import (CpuPlanner, O_TAINT, scene_c1_boxes, scene_c3_boxes)  # noqa: E402

NONE = np.zeros((0, 7))


def fresh(kind, obs, car=(0, 0, 0, 0, 0, 0), goal=(50, 0, 0, 0), seed=1):
    p = CpuPlanner(kind)
    p.set_obstacles(obs)
    p.srand(seed)
    p.tree_init(car, goal, 5.0)
    return p


def rollouts_both(tree, obs, parent, samples, gb, car, goal):
    """defined-variant outputs + mask of rows where the unmodified reference differs."""
    outs = []
    for kind in ("ref_defined", "ref"):
        p = fresh(kind, obs, car, goal)
        p.tree_import(tree)
        outs.append(p.rollout_batch(parent, samples, gb))
    d, u = outs
    differs = (d != u).any(axis=1) & ~(np.isnan(d) & np.isnan(u)).all(axis=1)
    tainted = d[:, O_TAINT] != 0
    assert not (differs & ~tainted).any(), "unmodified reference differs on an untainted rollout"
    return d, differs


def g0():
    p = CpuPlanner("ref")
    kat = {
        "obb_free": p.obb_dist([0, 0, 2, 5, 0], [3.55, 0, 2, 5, 1.57]),   # rrt/src/testers.cpp:132-133,158
        "obb_hit": p.obb_dist([0, 0, 2, 5, 0], [3.55, 0, 2, 5, 2]),       # rrt/src/testers.cpp:134,159
        "dubins_5_5_fwd": float(p.dubins(5, 5, 0, 0, 0, 1)),
        "dubins_5_5_rev": float(p.dubins(5, 5, 0, 0, 0, -1)),
        "dubins_1_1_in_circle": float(p.dubins(1, 1, 0, 0, 0, 1)),
    }
    p.set_obstacles([[10, 0, 0, 2, 4, 0, 0]])
    xs = [0, 4, 5, 5.5, 6, 7, 8, 9, 10, 11, 12, 13, 13.5, 14, 16]
    kat["vehicle_sweep_x"] = xs
    kat["vehicle_sweep_dobs"] = [p.obs_distance([x, 0, 0, 0, 0, 0, 0, 0, 0, 0]) for x in xs]
    p.set_obstacles(NONE)
    p.srand(1)
    p.tree_init()
    s, h, r = p.draw_samples(3)
    kat["first_samples"] = s.tolist()
    kat["first_heuristic"] = h.tolist()
    kat["first_r"] = r.tolist()
    kat["prius"] = p.vehicle().tolist()
    with open(os.path.join(HERE, "g0_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("G0", kat["obb_free"], kat["dubins_5_5_fwd"], kat["vehicle_sweep_dobs"][:4])


def g1_g2():
    car, goal = (0, 0, 0, 0, 3, 0), (50, 0, 0, 0)
    boxes = scene_c1_boxes()
    # tree snapshots: one tree grown by the reference with the 10 boxes; snapshot N = its first N nodes
    p = fresh("ref_defined", boxes, car, goal)
    p.expand(60)
    n60 = p.tree_size()
    while p.tree_size() < 1500:
        p.expand(50)
    tree = p.tree_export()[:1500]
    samples, heur, _ = p.draw_samples(4096)
    # ---- G2: candidate lists for snapshots of 1, 64, 250, 1500 nodes --------------------------------
    g2 = {"tree": tree, "samples": samples[:1024], "heuristic": heur[:1024], "n60": n60}
    for N in (1, 64, 250, 1500):
        q = fresh("ref_defined", boxes, car, goal)
        q.tree_import(tree[:N])
        cand, key, cnt = q.nearest_batch(samples[:1024], heur[:1024])
        g2[f"cand_{N}"], g2[f"key_{N}"], g2[f"count_{N}"] = cand, key, cnt
    np.savez_compressed(os.path.join(HERE, "g2_nearest.npz"), **g2)
    print("G2 tree", tree.shape, "n60", n60, "counts", [int(g2[f"count_{N}"].sum()) for N in (1, 64, 250, 1500)])
    # ---- G1: rollouts against the 60-iteration snapshot ----------------------------------------------
    snap = tree[:n60]
    q = fresh("ref_defined", boxes, car, goal)
    q.tree_import(snap)
    cand, key, cnt = q.nearest_batch(samples, heur)
    par, smp, rank = [], [], []
    for j in range(len(samples)):          # all (sample, candidate) pairs in reference order, first 4096
        for r in range(cnt[j]):
            par.append(cand[j, r]); smp.append(samples[j]); rank.append(r)
    par, smp, rank = np.array(par[:4096], np.int32), np.array(smp[:4096]), np.array(rank[:4096], np.int32)
    gb = np.zeros(len(par), np.uint8)
    # root-only subset (parent 0) and goal-biased subset (every node as parent)
    par_root = np.zeros(512, np.int32)
    smp_root = samples[-512:]
    par_gb = np.arange(n60, dtype=np.int32)
    g1 = {"tree": snap, "car": np.array(car, float), "goal": np.array(goal, float), "obstacles": boxes}
    for name, obs in (("obs", boxes), ("live", NONE)):
        o, dif = rollouts_both(snap, obs, par, smp, gb, car, goal)
        g1[f"out_{name}"], g1[f"unmod_differs_{name}"] = o, dif
        o, dif = rollouts_both(snap, obs, par_root, smp_root, np.zeros(512, np.uint8), car, goal)
        g1[f"out_root_{name}"], g1[f"unmod_differs_root_{name}"] = o, dif
        o, dif = rollouts_both(snap, obs, par_gb, np.zeros((n60, 2)), np.ones(n60, np.uint8), car, goal)
        g1[f"out_gb_{name}"], g1[f"unmod_differs_gb_{name}"] = o, dif
        print("G1", name, "tainted", int(g1[f"out_{name}"][:, O_TAINT].sum()), "of", len(par),
              "accepted", int(((g1[f"out_{name}"][:, 12] + g1[f"out_{name}"][:, 13]) > 0).sum()),
              "unmod differs", int(g1[f"unmod_differs_{name}"].sum()),
              "fail codes", np.bincount(g1[f"out_{name}"][:, 15].astype(int), minlength=4).tolist())
    g1.update(parent=par, samples=smp, rank=rank, parent_root=par_root, samples_root=smp_root, parent_gb=par_gb)
    np.savez_compressed(os.path.join(HERE, "g1_rollouts.npz"), **g1)


def g3():
    out = {}
    for name, obs in (("live", NONE), ("obs", scene_c1_boxes())):
        p = fresh("ref_defined", obs)
        s, h, _ = p.draw_samples(200)       # the draws expandTree will make (3 rand() per iteration)
        p.srand(1)
        p.expand(200)
        out[f"tree_{name}"] = p.tree_export()
        out[f"counters_{name}"] = np.array(list(p.counters().values()))
        out[f"best_{name}"] = p.best_path()
        out["samples"], out["heuristic"] = s, h
        u = fresh("ref", obs)
        u.expand(1)
        # node 1 of the live tree is the value quoted in SURVEY.md §8c (G3 row)
        print("G3", name, "nodes", p.tree_size(), "node1", out[f"tree_{name}"][1, :3], p.counters())
    np.savez_compressed(os.path.join(HERE, "g3_replay.npz"), **out)


def g4():
    car, goal = (0, 0, 0, 0, 3, 0), (100, 0, 0, 0)
    boxes = scene_c3_boxes()
    p = fresh("ref_defined", boxes, car, goal)
    p.expand(400)
    tree = p.tree_export()
    samples, heur, _ = p.draw_samples(1024)
    cand, key, cnt = p.nearest_batch(samples, heur)
    ok = cnt > 0
    par, smp = cand[ok, 0].astype(np.int32), samples[ok]
    o, dif = rollouts_both(tree, boxes, par, smp, np.zeros(len(par), np.uint8), car, goal)
    print("G4 tree", len(tree), "rollouts", len(par), "fail codes", np.bincount(o[:, 15].astype(int), minlength=4).tolist(),
          "tainted", int(o[:, O_TAINT].sum()))
    np.savez_compressed(os.path.join(HERE, "g4_dense.npz"), tree=tree, car=np.array(car, float), goal=np.array(goal, float),
                        obstacles=boxes, parent=par, samples=smp, out=o, unmod_differs=dif)


def g6():
    """Curved-road mode (MotionRequest.bend = true): the same rollouts as G1's obstacle scene with the
    lane-deviation cost of rrt/src/simulation.cpp:92-95 switched on (getDistToLane with laneShifts[0], Cxy)."""
    g1 = np.load(os.path.join(HERE, "g1_rollouts.npz"))
    Cxy, shift = (0.001, 0.02, 0.5), 1.5
    outs = []
    for kind in ("ref_defined", "ref"):
        p = CpuPlanner(kind)
        p.set_obstacles(g1["obstacles"])
        p.set_road(True, Cxy, shift)
        p.srand(1)
        p.tree_init(g1["car"], g1["goal"], 5.0)
        p.tree_import(g1["tree"])
        outs.append(p.rollout_batch(g1["parent"], g1["samples"], np.zeros(len(g1["parent"]), np.uint8)))
        p.set_road(False)
    d, u = outs
    differs = (d != u).any(axis=1) & ~(np.isnan(d) & np.isnan(u)).all(axis=1)
    assert not (differs & ~(d[:, O_TAINT] != 0)).any()
    base = g1["out_obs"]
    same_state = np.array_equal(d[:, :10], base[:, :10])
    print("G6 rollouts", len(d), "states equal to G1:", same_state, "mean extra costS", float(np.nanmean(d[:, 11] - base[:, 11])))
    assert same_state and (d[:, 11] >= base[:, 11]).all()
    np.savez_compressed(os.path.join(HERE, "g6_bend.npz"), Cxy=np.array(Cxy), lane_shift=np.array(shift), out=d,
                        unmod_differs=differs)


def g5(queries=100, iters=100):
    """Config C5: consecutive planMotion queries with commit_path = true (tree initialised from the previous best
    path, rrt/src/rrtplanner.cpp:50-94), moving obstacles, `iters` expandTree calls per query (tests/c5_scenario.py).
    Per query: the inputs (world state, goal and obstacles in the car frame) and the reference's results."""
    import c5_scenario as sc
    ref = CpuPlanner("ref_defined")
    ref.srand(1)
    ref.commit_reset()
    w = np.array([0, 0, 0, 0, 2.0, 0])
    t = 0.0
    out = dict(world=[], goal=[], obstacles=[], sizes=[], cost=[], sim_steps=[], nodes=[], n_nodes=[], rows=[])
    for q in range(queries):
        goal, obs = sc.to_car_frame(w, sc.world_goal(q), sc.world_obstacles(t))
        ref.set_obstacles(obs)
        carried, tree, nbest, steps, cost = ref.query_commit(w, goal, 5.0, iters)
        assert nbest > 0, f"query {q}: the reference found no path"
        nodes = ref.best_nodes(16)
        tr, rows = ref.best_traj()
        out["world"].append(w.copy()); out["goal"].append(goal); out["obstacles"].append(obs)
        out["sizes"].append([carried, tree, nbest]); out["cost"].append(cost); out["sim_steps"].append(steps)
        pad = np.zeros((16, nodes.shape[1])); pad[:len(nodes)] = nodes
        out["nodes"].append(pad); out["n_nodes"].append(len(nodes)); out["rows"].append(rows[:16].copy())
        w = sc.advance(w, tr, rows[:nbest])
        t += sc.DT_QUERY
    out = {k: np.array(v) for k, v in out.items()}
    out["iters"] = np.array(iters)
    print("G5", queries, "queries; carried", out["sizes"][:, 0].min(), "-", out["sizes"][:, 0].max(), "tree",
          out["sizes"][:, 1].min(), "-", out["sizes"][:, 1].max(), "car x final", out["world"][-1][0])
    np.savez_compressed(os.path.join(HERE, "g5_replan.npz"), **out)


def g7():
    """Curved-road mode on the host: the reference's road-frame transforms
    (rrt/src/transformations.cpp:20-202) and sampleOnLane (rrt/src/rrtplanner.cpp:204-224) on fixed inputs."""
    import ctypes as C
    ref = C.CDLL(os.path.join(os.path.dirname(HERE), "..", "oracle", "_ref", "libclrrt_ref.so"))
    ref.ref_init()
    ref.ref_road_transform.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    ref.ref_sample_on_lane.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    Cxy = np.array([-0.0035, 0.08, -1.0])
    x = np.linspace(0, 80, 400)
    dy = 2 * Cxy[0] * x + Cxy[1]
    S = np.concatenate([[0], np.cumsum(np.sqrt(1 + ((dy[1:] + dy[:-1]) / 2) ** 2) * np.diff(x))])
    Cxs = np.ascontiguousarray(np.polyfit(x, S, 2))
    rng = np.random.default_rng(21)
    n = 512
    px = rng.uniform(2, 70, n)
    inputs = np.column_stack([px, Cxy[0] * px ** 2 + Cxy[1] * px + Cxy[2] + rng.uniform(-6, 6, n), rng.uniform(-0.6, 0.6, n),
                              rng.uniform(-0.3, 0.3, n)])
    out = {"Cxy": Cxy, "Cxs": Cxs, "inputs": inputs}
    for what in range(7):
        r = inputs.copy()
        ref.ref_road_transform(what, Cxy.ctypes.data, Cxs.ctypes.data, r.ctypes.data, n)
        out[f"out{what}"] = r
    lanes = np.array([-3.5, 0.0, 3.5])
    K = 256
    s, h = np.zeros((K, 2)), np.zeros(K, np.uint8)
    C.CDLL(None).srand(C.c_uint(7))
    ref.ref_sample_on_lane(Cxy.ctypes.data, lanes.ctypes.data, 3, 60.0, 4.0, K, s.ctypes.data, h.ctypes.data)
    out.update(lanes=lanes, lane_samples=s, lane_heuristic=h)
    np.savez_compressed(os.path.join(HERE, "g7_road.npz"), **out)
    print("g7: road-frame transforms x 7,", K, "lane samples")


if __name__ == "__main__":
    if "--only-g7" in sys.argv:
        g7()
        sys.exit(0)
    if "--only-g5" in sys.argv:
        g5()
        sys.exit(0)
    if "--only-g6" in sys.argv:
        g6()
        sys.exit(0)
    g0(); g1_g2(); g3(); g4(); g5(); g6(); g7()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
