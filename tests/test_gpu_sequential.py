"""clrrt_expand_sequential: the reference's sequential expandTree loop (rrt/src/motionplanner.cpp:39-43) executed as
speculative windows on the device.  For every window size the tree, the parents, the counters and the best path must be
exactly those of the one-sample-at-a-time loop: the golden whole-query replay G3 (recorded from the reference's sources)
and fresh oracle runs with moving obstacles and a carried-over tree."""
import os

import numpy as np
import pytest

from cpulib import CpuPlanner, scene_c1_boxes
from gpu_common import clrrt, rel_err  # noqa: F401

pytestmark = pytest.mark.gpu
NONE = np.zeros((0, 7))
DISC_NODE = [7, 17, 18, 19]


@pytest.fixture(scope="module")
def planner(clrrt):
    pl = clrrt.Planner(device=0, tree_capacity=1 << 14, max_round=256)
    yield pl
    pl.close()


@pytest.mark.parametrize("window", [1, 3, 16, 64, 0])
@pytest.mark.parametrize("name", ["live", "obs"])
def test_g3_replay_in_windows(clrrt, planner, golden_dir, name, window):
    g = np.load(os.path.join(golden_dir, "g3_replay.npz"))
    planner.set_query((0, 0, 0, 0, 0, 0), (50, 0, 0, 0), 5.0)
    planner.set_obstacles(scene_c1_boxes() if name == "obs" else NONE)
    planner.tree_reset(clrrt.root_node((0, 0, 0, 0, 0, 0)))
    st = planner.expand_sequential(g["samples"], g["heuristic"], window=window)
    got, want = planner.tree_download_records(), g[f"tree_{name}"]
    assert st.iterations == 200 and st.tree_size == len(want) == len(got)
    assert np.array_equal(got[:, DISC_NODE], want[:, DISC_NODE])
    assert rel_err(got, want).max() == 0.0
    c = planner.counters()
    assert [c["fail_collision"], c["fail_acclimit"], c["fail_iterlimit"], c["sim_count"]] == g[f"counters_{name}"].tolist()
    assert st.sim_steps == c["sim_count"]
    assert np.array_equal(planner.best_path(), g[f"best_{name}"])
    if window == 1:
        assert st.windows + st.exact_fallbacks >= 200
    print(f"{name} window {window}: {st.windows} windows for 200 iterations ({st.speculated} samples speculated, "
          f"{st.exact_fallbacks} via the std::sort route), {st.ms_total:.1f} ms")


@pytest.mark.parametrize("seed,moving", [(3, False), (4, True)])
def test_sequential_windows_equal_the_oracle_loop(clrrt, planner, seed, moving):
    """600 iterations from a tree the oracle has grown, in two calls (the second continues the first)."""
    car, goal = (0, 0, 0, 0, 2, 0), (55, 1.5, 0.05, 0)
    obs = scene_c1_boxes(moving=moving)
    orc = CpuPlanner("oracle")
    orc.set_tie_mode(1)
    orc.set_obstacles(obs)
    orc.srand(seed)
    orc.tree_init(car, goal, 5.0)
    orc.expand(30)
    planner.set_query(car, goal, 5.0)
    planner.set_obstacles(obs)
    planner.tree_reset_records(orc.tree_export())
    c0 = orc.counters()
    s, h, _ = orc.draw_samples(600)
    orc.expand_with(s, h)
    st1 = planner.expand_sequential(s[:250], h[:250], window=0)
    st2 = planner.expand_sequential(s[250:], h[250:], window=24)
    a, b = planner.tree_download_records(), orc.tree_export()
    assert len(a) == len(b) == st2.tree_size
    assert np.array_equal(a[:, DISC_NODE], b[:, DISC_NODE])
    assert rel_err(a, b).max() == 0.0
    c1, gc = orc.counters(), planner.counters()
    keys = ["fail_collision", "fail_acclimit", "fail_iterlimit", "sim_count"]
    assert [gc[k] for k in keys] == [int(c1[k] - c0[k]) for k in keys]
    assert st1.sim_steps + st2.sim_steps == gc["sim_count"]
    assert st1.windows + st2.windows < 400, "speculation should commit more than one sample per window on average"


def test_long_query_with_equal_keys_equals_the_oracle_loop(clrrt, planner):
    """1500 iterations of the C1 query from the root.  With the optimise key the sums costE + length of the nodes along the
    straight line from the root coincide for one sample in ten: the three routes for equal keys — they cannot change the
    outcome (committed as speculated), the reference's own sort gives the same outcome (commit repeated), it does not (K = 1
    host route) — must all occur here, and the tree must be the oracle's with std::sort's order of equal keys."""
    car, goal = (0, 0, 0, 0, 0, 0), (50, 0, 0, 0)
    obs = scene_c1_boxes()
    orc = CpuPlanner("oracle")
    orc.set_tie_mode(1)
    orc.set_obstacles(obs)
    orc.srand(1)
    orc.tree_init(car, goal, 5.0)
    planner.set_query(car, goal, 5.0)
    planner.set_obstacles(obs)
    planner.tree_reset_records(orc.tree_export())
    c0 = orc.counters()
    s, h, _ = orc.draw_samples(1500)
    orc.expand_with(s, h)
    st = planner.expand_sequential(s, h, window=0)
    a, b = planner.tree_download_records(), orc.tree_export()
    assert len(a) == len(b) == st.tree_size
    assert np.array_equal(a[:, DISC_NODE], b[:, DISC_NODE])
    assert rel_err(a, b).max() == 0.0
    c1, gc = orc.counters(), planner.counters()
    keys = ["fail_collision", "fail_acclimit", "fail_iterlimit", "sim_count"]
    assert [gc[k] for k in keys] == [int(c1[k] - c0[k]) for k in keys]
    print(f"1500 iterations: {st.windows} windows, {st.tie_checks_same} equal-key samples with the same outcome under the "
          f"reference's sort, {st.exact_fallbacks} through the K = 1 route")
    assert st.tie_checks_same > 0 and st.exact_fallbacks > 0
    assert st.windows < 500


def test_sequential_stops_at_a_full_tree(clrrt):
    pl = clrrt.Planner(device=0, tree_capacity=40, max_round=64)
    pl.set_query((0, 0, 0, 0, 0, 0), (50, 0, 0, 0), 5.0)
    pl.set_obstacles(NONE)
    pl.tree_reset(clrrt.root_node((0, 0, 0, 0, 0, 0)))
    s, h = clrrt.draw_samples((50, 0, 0, 0), 200, seed=1)
    with pytest.raises(clrrt.ClrrtError, match="capacity"):
        pl.expand_sequential(s, h, window=16)
    n = pl.tree_size()
    assert 38 <= n <= 40
    # what was appended is the prefix of the sequential tree
    big = clrrt.Planner(device=0, tree_capacity=4096, max_round=64)
    big.set_query((0, 0, 0, 0, 0, 0), (50, 0, 0, 0), 5.0)
    big.set_obstacles(NONE)
    big.tree_reset(clrrt.root_node((0, 0, 0, 0, 0, 0)))
    big.expand_sequential(s, h, window=1)
    assert pl.tree_download().tobytes() == big.tree_download()[:n].tobytes()
    pl.close(); big.close()
