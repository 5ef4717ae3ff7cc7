"""The C restatement (oracle/clrrt_oracle.c) against the golden vectors generated from the reference's own
sources (tests/golden/make_golden.py).  Bit-for-bit: both sides ran on glibc libm without FMA contraction."""
import json
import os

import numpy as np
import pytest

from cpulib import CpuPlanner, O_TAINT, check_candidate_lists, scene_c1_boxes

NONE = np.zeros((0, 7))


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.fixture(scope="module")
def orc():
    return CpuPlanner("oracle")


def test_g0_known_answers(orc, golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "g0_kat.json")))
    # rrt/src/testers.cpp:132-134,158-159 — the reference's own (commented-out) collision assertions
    free = orc.obb_dist([0, 0, 2, 5, 0], [3.55, 0, 2, 5, 1.57])
    hit = orc.obb_dist([0, 0, 2, 5, 0], [3.55, 0, 2, 5, 2])
    assert free > 0 and hit == 0
    assert free == kat["obb_free"] and abs(free - 0.096019268) < 1e-8
    assert float(orc.dubins(5, 5, 0, 0, 0, 1)) == kat["dubins_5_5_fwd"]
    assert float(orc.dubins(5, 5, 0, 0, 0, -1)) == kat["dubins_5_5_rev"]
    assert float(orc.dubins(1, 1, 0, 0, 0, 1)) == kat["dubins_1_1_in_circle"]
    orc.set_obstacles([[10, 0, 0, 2, 4, 0, 0]])
    got = [orc.obs_distance([x, 0, 0, 0, 0, 0, 0, 0, 0, 0]) for x in kat["vehicle_sweep_x"]]
    assert got == kat["vehicle_sweep_dobs"]
    orc.set_obstacles(NONE)
    orc.srand(1)
    orc.tree_init()
    s, h, r = orc.draw_samples(3)
    assert s.tolist() == kat["first_samples"] and h.tolist() == kat["first_heuristic"] and r.tolist() == kat["first_r"]
    assert abs(s[0, 0] - 50.41126251220703) < 1e-12 and abs(r[0] - 0.7830992237586059) < 1e-15
    assert orc.vehicle().tolist() == kat["prius"]


@pytest.mark.parametrize("name", ["obs", "live"])
def test_g1_rollouts(orc, golden_dir, name):
    g = np.load(os.path.join(golden_dir, "g1_rollouts.npz"))
    orc.set_obstacles(g["obstacles"] if name == "obs" else NONE)
    orc.tree_init(g["car"], g["goal"], 5.0)
    orc.tree_import(g["tree"])
    out = orc.rollout_batch(g["parent"], g["samples"])
    assert same(out, g[f"out_{name}"])
    out = orc.rollout_batch(g["parent_root"], g["samples_root"])
    assert same(out, g[f"out_root_{name}"])
    n = len(g["parent_gb"])
    out = orc.rollout_batch(g["parent_gb"], np.zeros((n, 2)), np.ones(n, np.uint8))
    assert same(out, g[f"out_gb_{name}"])
    # the unmodified reference may only differ on UB-tainted rollouts
    assert not (g[f"unmod_differs_{name}"] & (g[f"out_{name}"][:, O_TAINT] == 0)).any()


def test_g6_curved_road_cost(orc, golden_dir):
    """bend = true: getDistToLane in every step's cost (rrt/src/simulation.cpp:49-53, :92-95), bit for bit."""
    g = np.load(os.path.join(golden_dir, "g1_rollouts.npz"))
    g6 = np.load(os.path.join(golden_dir, "g6_bend.npz"))
    orc.set_obstacles(g["obstacles"])
    orc.set_road(True, g6["Cxy"], float(g6["lane_shift"]))
    try:
        orc.tree_init(g["car"], g["goal"], 5.0)
        orc.tree_import(g["tree"])
        out = orc.rollout_batch(g["parent"], g["samples"])
        assert same(out, g6["out"])
    finally:
        orc.set_road(False)


def test_g2_candidate_lists(orc, golden_dir):
    g = np.load(os.path.join(golden_dir, "g2_nearest.npz"))
    orc.set_obstacles(scene_c1_boxes())
    orc.tree_init((0, 0, 0, 0, 3, 0), (50, 0, 0, 0), 5.0)
    for N in (1, 64, 250, 1500):
        orc.tree_import(g["tree"][:N])
        cand, key, cnt = orc.nearest_batch(g["samples"], g["heuristic"])
        assert same(cnt, g[f"count_{N}"])
        assert same(key, g[f"key_{N}"])
        # Node order may legitimately differ between EQUAL keys only (std::sort is unstable upstream, and the
        # reference's trees hold many one-step nodes with identical poses, hence identical keys): the golden
        # lists must be valid answers under the oracle's recomputed keys, and where no tie is involved the ids match.
        check_candidate_lists(orc, g["samples"], g["heuristic"], g[f"cand_{N}"], g[f"key_{N}"], g[f"count_{N}"])
        check_candidate_lists(orc, g["samples"], g["heuristic"], cand, key, cnt)
        untied = np.ones_like(cand, bool)
        untied[:, 1:] &= key[:, 1:] != key[:, :-1]
        untied[:, :-1] &= key[:, :-1] != key[:, 1:]
        untied[:, -1] = False  # a tie may continue past the cut at rank 10
        assert same(cand[untied], g[f"cand_{N}"][untied])


@pytest.mark.parametrize("name", ["live", "obs"])
def test_g3_whole_query_replay(orc, golden_dir, name):
    g = np.load(os.path.join(golden_dir, "g3_replay.npz"))
    orc.set_obstacles(scene_c1_boxes() if name == "obs" else NONE)
    orc.srand(1)
    orc.tree_init()
    s, h, _ = orc.draw_samples(200)
    assert same(s, g["samples"]) and same(h, g["heuristic"])
    orc.srand(1)
    orc.expand(200)
    assert same(orc.tree_export(), g[f"tree_{name}"])
    assert list(orc.counters().values()) == g[f"counters_{name}"].tolist()
    assert same(orc.best_path(), g[f"best_{name}"])
    if name == "live":  # the node quoted in SURVEY.md §8c
        t = orc.tree_export()
        assert t[1, 0] == 46.723439755885778 and t[1, 7] == 245
    # the same tree again through the caller-supplied-sample entry point
    orc.tree_init()
    orc._f("expand_with")(s.ctypes.data, h.ctypes.data, 200)
    assert same(orc.tree_export(), g[f"tree_{name}"])


def test_g4_dense_scene(orc, golden_dir):
    g = np.load(os.path.join(golden_dir, "g4_dense.npz"))
    orc.set_obstacles(g["obstacles"])
    orc.tree_init(g["car"], g["goal"], 5.0)
    orc.tree_import(g["tree"])
    out = orc.rollout_batch(g["parent"], g["samples"])
    assert same(out, g["out"])
