"""clrrt_simulate — the reference's Simulation with its own parameter list (rrt/include/rrt/simulation.h:18-19): arbitrary
start state, caller-owned reference path of any shape, GoalBiased / genProfile / Vstart — against the oracle's restatement
and, when oracle/_ref is present, the reference's own constructor.  Discrete outputs exact, states / costs / the filled
ref.v bit-equal."""
import numpy as np
import pytest

from cpulib import CpuPlanner, ref_available, scene_c1_boxes
from gpu_common import assert_rollouts_match, clrrt  # noqa: F401

pytestmark = pytest.mark.gpu
GOAL = (50, 0, 0, 0)


def make_cases(seed, n):
    """References of several shapes (none of them the straight, equally spaced lines expandTree builds) with start states
    on and off the path."""
    rng = np.random.default_rng(seed)
    cases = []
    for k in range(n):
        shape = k % 5
        npts = int(rng.integers(3, 400))
        s = np.linspace(0, rng.uniform(1, 60), npts)
        if shape == 0:    # sinusoid
            x, y = s, rng.uniform(0.2, 3) * np.sin(s / rng.uniform(3, 15))
        elif shape == 1:  # arc
            R = rng.uniform(8, 80) * rng.choice([-1, 1])
            x, y = R * np.sin(s / R), R * (1 - np.cos(s / R))
        elif shape == 2:  # polyline with a kink and unequal spacing
            x = np.cumsum(rng.uniform(0.05, 0.5, npts))
            y = np.where(x > x[npts // 2], (x - x[npts // 2]) * rng.uniform(-0.5, 0.5), 0.0)
        elif shape == 3:  # straight, but arbitrary orientation / offset
            a = rng.uniform(-0.6, 0.6)
            x, y = s * np.cos(a) + rng.uniform(-2, 2), s * np.sin(a) + rng.uniform(-2, 2)
        else:             # S-curve ending near the goal
            x = np.linspace(0, 50, npts)
            y = 2.5 * np.tanh((x - 25) / 6) * rng.uniform(-1, 1)
        st = np.zeros(10)
        st[0] = x[0] + rng.uniform(-1, 1); st[1] = y[0] + rng.uniform(-1, 1)
        st[2] = np.arctan2(y[min(2, npts - 1)] - y[0], x[min(2, npts - 1)] - x[0]) + rng.uniform(-0.3, 0.3)
        st[3] = rng.uniform(-0.2, 0.2); st[4] = rng.uniform(0, 5); st[5] = rng.uniform(-1, 1); st[6] = rng.uniform(0, 10)
        st[8] = rng.uniform(0, 5); st[9] = rng.uniform(-0.3, 0.3)
        gen = k % 7 != 3
        v = None if gen else np.clip(3 + np.sin(np.arange(npts) / 9.0), 0, 5)
        cases.append(dict(state=st, x=np.ascontiguousarray(x), y=np.ascontiguousarray(y), v=v, gb=int(k % 4 == 1), gen=int(gen),
                          vstart=float(rng.uniform(0, 5)), dir=1 if k % 11 else -1))
    return cases


def cpu_run(cpu, c):
    out, v, _ = cpu.simulate(c["state"], c["x"], c["y"], c["v"], gb=c["gb"], gen_profile=c["gen"], vstart=c["vstart"], ref_dir=c["dir"])
    return out, v


@pytest.mark.parametrize("scene", ["free", "boxes", "moving"])
def test_simulate_matches_the_reference_constructor(clrrt, scene):
    obs = {"free": np.zeros((0, 7)), "boxes": scene_c1_boxes(), "moving": scene_c1_boxes(moving=True)}[scene]
    cases = make_cases({"free": 1, "boxes": 2, "moving": 3}[scene], 300)
    orc = CpuPlanner("oracle")
    orc.set_obstacles(obs)
    orc.tree_init((0, 0, 0, 0, 2, 0), GOAL, 5.0)
    want = [cpu_run(orc, c) for c in cases]
    pl = clrrt.Planner(device=0, tree_capacity=64, max_round=64)
    pl.set_query((0, 0, 0, 0, 2, 0), GOAL, 5.0)
    pl.set_obstacles(obs)
    c0 = pl.counters()
    # one by one through clrrt_simulate (the constructor's shape) ...
    got = [pl.simulate(c["state"], c["x"], c["y"], c["v"], goal_biased=c["gb"], gen_profile=c["gen"], vstart=c["vstart"],
                       ref_dir=c["dir"]) for c in cases[:60]]
    tab = clrrt.rollouts_as_table(np.array([g[0] for g in got]))
    assert_rollouts_match(tab, np.array([w[0] for w in want[:60]]), f"clrrt_simulate, {scene}")
    for g, w in zip(got, want):
        assert np.array_equal(g[1], w[1]), "ref.v filled by generateVelocityProfile differs"
    # ... and all of them in one launch (ragged references)
    rec, vs = pl.simulate_batch([c["state"] for c in cases], [(c["x"], c["y"], c["v"]) for c in cases],
                                goal_biased=[c["gb"] for c in cases], gen_profile=[c["gen"] for c in cases],
                                vstart=[c["vstart"] for c in cases], ref_dir=[c["dir"] for c in cases])
    tab = clrrt.rollouts_as_table(rec)
    wtab = np.array([w[0] for w in want])
    assert_rollouts_match(tab, wtab, f"clrrt_simulate_batch, {scene}")
    for v, w in zip(vs, want):
        assert np.array_equal(v, w[1])
    # the global counters advance as upstream (rrt/src/simulation.cpp:59, :85, :102, :142)
    c1 = pl.counters()
    n = len(cases) + 60
    assert c1["rollouts"] - c0["rollouts"] == n
    assert c1["sim_count"] - c0["sim_count"] == int(wtab[:, 14].sum() + wtab[:60, 14].sum())
    assert c1["fail_collision"] - c0["fail_collision"] == int((wtab[:, 15] == 1).sum() + (wtab[:60, 15] == 1).sum())
    codes = np.bincount(wtab[:, 15].astype(int), minlength=4)
    assert (wtab[:, 12] == 1).sum() > 20 and codes[2] + codes[3] > 0, "the cases should end in several different ways"
    if scene != "free":
        assert codes[1] > 5
    # trajectories
    c = cases[0]
    r, v, traj = pl.simulate(c["state"], c["x"], c["y"], c["v"], goal_biased=c["gb"], gen_profile=c["gen"], vstart=c["vstart"],
                             ref_dir=c["dir"], traj_stride=512)
    _, _, wtraj = orc.simulate(c["state"], c["x"], c["y"], c["v"], gb=c["gb"], gen_profile=c["gen"], vstart=c["vstart"], ref_dir=c["dir"])
    assert np.array_equal(traj[:len(wtraj)], wtraj), "stateArray differs"
    pl.close()


@pytest.mark.skipif(not ref_available(True), reason="oracle/_ref not built")
def test_simulate_matches_the_reference_binary(clrrt):
    """Same, against the reference's own Simulation constructor (index-clamped build)."""
    cases = make_cases(9, 120)
    ref = CpuPlanner("ref_defined")
    ref.set_obstacles(scene_c1_boxes())
    ref.tree_init((0, 0, 0, 0, 2, 0), GOAL, 5.0)
    want = [cpu_run(ref, c) for c in cases]
    pl = clrrt.Planner(device=0, tree_capacity=64, max_round=64)
    pl.set_query((0, 0, 0, 0, 2, 0), GOAL, 5.0)
    pl.set_obstacles(scene_c1_boxes())
    rec, vs = pl.simulate_batch([c["state"] for c in cases], [(c["x"], c["y"], c["v"]) for c in cases],
                                goal_biased=[c["gb"] for c in cases], gen_profile=[c["gen"] for c in cases],
                                vstart=[c["vstart"] for c in cases], ref_dir=[c["dir"] for c in cases])
    assert_rollouts_match(clrrt.rollouts_as_table(rec), np.array([w[0] for w in want]), "clrrt_simulate_batch vs oracle/_ref")
    for v, w in zip(vs, want):
        assert np.array_equal(v, w[1])
    pl.close()


def test_simulate_argument_errors(clrrt):
    pl = clrrt.Planner(device=0, tree_capacity=64, max_round=64)
    with pytest.raises(clrrt.ClrrtError):
        pl.simulate(np.zeros(10), [0.0, 1.0], [0.0, 0.0])            # fewer than three reference points
    with pytest.raises(clrrt.ClrrtError):
        pl.simulate(np.zeros(10), [0.0, 1.0, 2.0], [0.0, 0.0, 0.0], ref_dir=0)
    pl.close()
