"""GPU parity of the rollout kernel (clrrt_propagate_batch) against the golden vectors generated from the
reference's own sources and against the oracle on fresh seeded inputs.  Discrete outputs (verdicts, step counts,
waypoint traces) bit-exact; states and costs within 1e-6 relative."""
import os

import numpy as np
import pytest

from cpulib import CpuPlanner, O_TAINT, scene_c1_boxes, scene_c3_boxes
from gpu_common import CONT, DISC, assert_rollouts_match, clrrt, rel_err  # noqa: F401

pytestmark = pytest.mark.gpu
NONE = np.zeros((0, 7))


@pytest.fixture(scope="module")
def g1(golden_dir):
    return np.load(os.path.join(golden_dir, "g1_rollouts.npz"))


@pytest.fixture(scope="module")
def planner(clrrt):
    pl = clrrt.Planner(device=0, tree_capacity=1 << 17, max_round=1 << 16)
    yield pl
    pl.close()


@pytest.mark.parametrize("name", ["obs", "live"])
def test_g1_rollouts_vs_reference_golden(clrrt, planner, g1, name):
    planner.set_query(g1["car"], g1["goal"], 5.0)
    planner.tree_reset_records(g1["tree"])
    planner.set_obstacles(g1["obstacles"] if name == "obs" else NONE)
    got = clrrt.rollouts_as_table(planner.propagate_batch(g1["parent"], g1["samples"]))
    assert_rollouts_match(got, g1[f"out_{name}"], f"G1 {name}")
    got = clrrt.rollouts_as_table(planner.propagate_batch(g1["parent_root"], g1["samples_root"]))
    assert_rollouts_match(got, g1[f"out_root_{name}"], f"G1 root {name}")
    # untainted rows are results of the UNMODIFIED reference; tainted rows of its "defined" variant
    tainted = g1[f"out_{name}"][:, O_TAINT] != 0
    assert not (g1[f"unmod_differs_{name}"] & ~tainted).any()
    print(f"tainted fraction {tainted.mean():.3f}")


def test_g6_curved_road_cost(clrrt, planner, g1, golden_dir):
    """bend = true: the lane-deviation term of rrt/src/simulation.cpp:92-95 in every step's cost; states and
    verdicts are those of G1, costS carries the extra term (golden G6 from the reference's sources)."""
    g6 = np.load(os.path.join(golden_dir, "g6_bend.npz"))
    planner.set_query(g1["car"], g1["goal"], 5.0)
    planner.set_road(True, g6["Cxy"], float(g6["lane_shift"]))
    try:
        planner.tree_reset_records(g1["tree"])
        planner.set_obstacles(g1["obstacles"])
        got = clrrt.rollouts_as_table(planner.propagate_batch(g1["parent"], g1["samples"]))
        assert_rollouts_match(got, g6["out"], "G6 bend")
        assert (got[:, 11] >= g1["out_obs"][:, 11]).all() and (got[:, 11] > g1["out_obs"][:, 11]).mean() > 0.9
    finally:
        planner.set_road(False)


@pytest.mark.parametrize("name", ["obs", "live"])
def test_g1_goal_biased_rollouts(clrrt, planner, g1, name):
    """Goal-biased references duplicate their junction point (rrt/src/reference.cpp:56-63): for the one or two steps
    whose 3-point window straddles the pair, the reference's Lagrange interpolation divides by ~1e-15 and the steer
    command is rounding noise that depends on the last bit of libm's sin/cos.  Those rollouts
    are therefore compared exactly up to the junction and with a loose bound after it; verdicts must still agree."""
    planner.set_query(g1["car"], g1["goal"], 5.0)
    planner.tree_reset_records(g1["tree"])
    planner.set_obstacles(g1["obstacles"] if name == "obs" else NONE)
    n = len(g1["parent_gb"])
    got = clrrt.rollouts_as_table(planner.propagate_batch(g1["parent_gb"], np.zeros((n, 2)), np.ones(n, np.uint8)))
    want = g1[f"out_gb_{name}"]
    verdict = [12, 13, 15, 16, 22]
    assert np.array_equal(got[:, verdict], want[:, verdict])
    assert np.array_equal(got[:, 17:20], want[:, 17:20])  # reference geometry and velocity profile: exact
    acc = (want[:, 12] + want[:, 13]) > 0                  # rollouts the planner keeps
    assert np.abs(got[acc, 14] - want[acc, 14]).max() <= 3  # step counts may move by the noisy steps
    assert np.abs(got[acc, :3] - want[acc, :3]).max() < 0.05  # final pose within 5 cm / 0.05 rad
    exact = (got[:, :7] == want[:, :7]).all(1)
    print(f"goal-biased rollouts bit-identical to the reference: {int(exact.sum())} of {n}")
    # with glibc's double sin/cos/tan restated on the device (csrc/refmath64.cuh) the noisy steps come out the same
    assert exact.all() and np.array_equal(got[acc, 14], want[acc, 14])


def test_g4_dense_scene_vs_reference_golden(clrrt, planner, golden_dir):
    g = np.load(os.path.join(golden_dir, "g4_dense.npz"))
    planner.set_query(g["car"], g["goal"], 5.0)
    planner.set_obstacles(g["obstacles"])
    planner.tree_reset_records(g["tree"])
    got = clrrt.rollouts_as_table(planner.propagate_batch(g["parent"], g["samples"]))
    assert_rollouts_match(got, g["out"], "G4")


@pytest.mark.parametrize("scene", ["static", "moving", "none"])
def test_fresh_batch_vs_oracle(clrrt, planner, scene):
    """20 000 seeded rollouts (all candidate ranks) against the oracle: sizes the oracle finishes in seconds."""
    obs = {"static": scene_c1_boxes(), "moving": scene_c1_boxes(moving=True), "none": NONE}[scene]
    car, goal = (0, 0, 0, 0, 2, 0), (50, 0, 0, 0)
    orc = CpuPlanner("oracle")
    orc.set_obstacles(obs)
    orc.srand(11)
    orc.tree_init(car, goal, 5.0)
    orc.expand(120)
    tree = orc.tree_export()
    s, h, _ = orc.draw_samples(3000)
    cand, key, cnt = orc.nearest_batch(s, h)
    par = np.concatenate([cand[j, :cnt[j]] for j in range(len(s))])[:20000]
    smp = np.concatenate([np.repeat(s[j:j + 1], cnt[j], 0) for j in range(len(s))])[:20000]
    want = orc.rollout_batch(par, smp)
    planner.set_query(car, goal, 5.0)
    planner.set_obstacles(obs)
    planner.tree_reset_records(tree)
    got = clrrt.rollouts_as_table(planner.propagate_batch(par, smp))
    e = assert_rollouts_match(got, want, scene)
    assert e == 0.0, "states and costs are expected bit-equal to the oracle's (csrc/refmath64.cuh)"
    print(f"{scene}: {len(par)} rollouts, fail codes {np.bincount(want[:, 15].astype(int), minlength=4)}, max rel err {e:.2e}")


def test_exact_distance_mode_vs_oracle(clrrt, planner):
    """weight_obstacle_gain != 0: the pseudo-distance returned by getOBBdist enters costS, so the kernel evaluates
    every obstacle's first separating axis as the reference does (no broad phase)."""
    obs = scene_c1_boxes(moving=True)
    obs[::3, 5] = 0.0  # mix of static and moving boxes
    car, goal = (0, 0, 0, 0, 2, 0), (50, 0, 0, 0)
    w = [10, 5, 2.0, 4, 1]
    orc = CpuPlanner("oracle")
    orc.set_weights(w)
    orc.set_obstacles(obs)
    orc.srand(21)
    orc.tree_init(car, goal, 5.0)
    orc.expand(80)
    tree = orc.tree_export()
    s, h, _ = orc.draw_samples(1500)
    cand, key, cnt = orc.nearest_batch(s, h)
    ok = cnt > 0
    par, smp = cand[ok, 0], s[ok]
    want = orc.rollout_batch(par, smp)
    p = clrrt.default_params()
    for i in range(5):
        p.Wcost[i] = w[i]
    pl = clrrt.Planner(params=p, device=0, tree_capacity=1 << 12, max_round=1 << 11)
    pl.set_query(car, goal, 5.0)
    pl.set_obstacles(obs)
    pl.tree_reset_records(tree)
    got = clrrt.rollouts_as_table(pl.propagate_batch(par, smp))
    assert_rollouts_match(got, want, "exact distance")  # costS holds exp(-W3 Dobs): glibc's exp restated (refmath64.cuh)
    assert (want[:, 11] > 0).any()
    pl.close()
    orc.set_weights([10, 5, 0, 4, 1])


def test_trajectories_match_oracle(clrrt, planner, g1):
    planner.set_query(g1["car"], g1["goal"], 5.0)
    planner.tree_reset_records(g1["tree"])
    planner.set_obstacles(g1["obstacles"])
    orc = CpuPlanner("oracle")
    orc.set_obstacles(g1["obstacles"])
    orc.tree_init(g1["car"], g1["goal"], 5.0)
    orc.tree_import(g1["tree"])
    idx = np.arange(0, 4096, 257)
    out, traj = planner.propagate_batch(g1["parent"][idx], g1["samples"][idx], traj_stride=512)
    for k, i in enumerate(idx):
        want, _ = orc.rollout_traj(int(g1["parent"][i]), g1["samples"][i], 0)
        n = out["n_steps"][k] + 1
        assert n == len(want)
        assert np.array_equal(traj[k, :n, 7], want[:, 7])
        assert rel_err(traj[k, :n], want).max() < 1e-6


def test_edge_cases(clrrt, planner, g1):
    planner.set_query(g1["car"], g1["goal"], 5.0)
    planner.tree_reset_records(g1["tree"])
    planner.set_obstacles(g1["obstacles"])
    one = clrrt.rollouts_as_table(planner.propagate_batch(g1["parent"][:1], g1["samples"][:1]))
    assert_rollouts_match(one, g1["out_obs"][:1], "M=1")
    with pytest.raises(clrrt.ClrrtError):
        planner.propagate_batch([len(g1["tree"])], [[1.0, 1.0]])  # parent out of range
    with pytest.raises(clrrt.ClrrtError):
        planner.propagate_batch([-1], [[1.0, 1.0]])
    # ragged batch sizes around the warp/block granularity give the same rows
    for m in (31, 33, 127, 129, 1000):
        got = clrrt.rollouts_as_table(planner.propagate_batch(g1["parent"][:m], g1["samples"][:m]))
        assert_rollouts_match(got, g1["out_obs"][:m], f"M={m}")


def test_full_size_properties(clrrt, planner):
    """Config-C3 size (65 536 rollouts, 1000 boxes): results do not depend on batch composition or lane placement
    (a rollout is a pure function of parent, sample and scene), and a random subset equals the oracle."""
    boxes = scene_c3_boxes()
    car, goal = (0, 0, 0, 0, 3, 0), (100, 0, 0, 0)
    orc = CpuPlanner("oracle")
    orc.set_obstacles(boxes)
    orc.srand(4)
    orc.tree_init(car, goal, 5.0)
    orc.expand(150)
    tree = orc.tree_export()
    s, h, _ = orc.draw_samples(65536)
    rng = np.random.default_rng(0)
    par = rng.integers(0, len(tree), 65536).astype(np.int32)
    planner.set_query(car, goal, 5.0)
    planner.set_obstacles(boxes)
    planner.tree_reset_records(tree)
    a = planner.propagate_batch(par, s)
    perm = rng.permutation(65536)
    b = planner.propagate_batch(par[perm], s[perm])
    assert a[perm].tobytes() == b.tobytes()
    sub = rng.choice(65536, 300, replace=False)
    want = orc.rollout_batch(par[sub], s[sub])
    ok = want[:, 16] >= 3  # random (parent, sample) pairs may give references shorter than 3 points: undefined upstream
    assert_rollouts_match(clrrt.rollouts_as_table(a[sub])[ok], want[ok], "C3 subset")
    # launch tuning (work-refill threshold, persistent-grid size, broad-phase cell size) must not change any result
    planner.set_tuning(refill_min=1, blocks_per_sm=1)
    planner.set_grid_cell(0.3)
    planner.set_obstacles(boxes)
    c = planner.propagate_batch(par, s)
    planner.set_grid_cell(-7.0)  # position grid only, cells with more than 64 listed obstacles: the chunked broad phase
    planner.set_obstacles(boxes)
    d = planner.propagate_batch(par, s)
    planner.set_tuning(refill_min=8, blocks_per_sm=0)
    planner.set_grid_cell(1.0)
    planner.set_obstacles(boxes)
    assert a.tobytes() == c.tobytes() and a.tobytes() == d.tobytes()


def test_branch_free_division(clrrt):
    """rollout.cuh's div_nb — the compiler's own fast path of the double division with its range test accumulated instead of
    branched on — gives the operator's result bit for bit wherever it accepts its own result: 2.7e8 operand pairs over the
    rollout's magnitudes, random bit patterns (every exponent, NaN, Inf, subnormals) and the edges of the accepted range."""
    import ctypes as C
    pl = clrrt.Planner(device=0, tree_capacity=64, max_round=64)
    lib = clrrt.load_library()
    lib.clrrt_debug_div_check.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, C.c_void_p]
    out = (C.c_ulonglong * 3)()
    assert lib.clrrt_debug_div_check(pl.h, 12345, 900, out) == 0
    accepted, differ, slow = out[0], out[1], out[2]
    assert accepted + slow == 900 * 148 * 8 * 256 or accepted + slow > 1e8
    assert differ == 0, f"{differ} of {accepted} fast-path quotients differ from a / b"
    assert accepted > 0.5 * (accepted + slow) and slow > 0
    pl.close()
