"""Edge cases of the obstacle handling and of the round bookkeeping on the GPU, against the oracle:
obstacle sets beyond the shared-memory table, pose-grid cells that overflow, scenes far from the origin, degenerate
boxes, many moving obstacles, non-finite samples, a full tree."""
import numpy as np
import pytest

from cpulib import CpuPlanner, scene_c1_boxes
from gpu_common import assert_rollouts_match, clrrt  # noqa: F401

pytestmark = pytest.mark.gpu
CAR, GOAL = (0, 0, 0, 0, 2, 0), (50, 0, 0, 0)


@pytest.fixture(scope="module")
def planner(clrrt):
    pl = clrrt.Planner(device=0, tree_capacity=1 << 14, max_round=1 << 12)
    yield pl
    pl.close()


def oracle_batch(obs, seed, grow=100, n=1500, car=CAR, goal=GOAL, tree_obs=None):
    """A tree grown by the oracle (among `tree_obs`, default the C1 boxes) and n top-1 rollouts among `obs`."""
    orc = CpuPlanner("oracle")
    orc.set_obstacles(scene_c1_boxes() if tree_obs is None else tree_obs)
    orc.srand(seed)
    orc.tree_init(car, goal, 5.0)
    orc.expand(grow)
    tree = orc.tree_export()
    s, h, _ = orc.draw_samples(n)
    cand, key, cnt = orc.nearest_batch(s, h)
    ok = cnt > 0
    par, smp = cand[ok, 0], s[ok]
    orc.set_obstacles(obs)
    want = orc.rollout_batch(par, smp)
    return tree, par, smp, want


def run(clrrt, planner, obs, tree, par, smp, car=CAR, goal=GOAL):
    planner.set_query(car, goal, 5.0)
    planner.set_obstacles(obs)
    planner.tree_reset_records(tree)
    return clrrt.rollouts_as_table(planner.propagate_batch(par, smp))


def test_obstacle_table_beyond_shared_memory(clrrt, planner):
    """3000 static boxes: the 32-byte broad-phase records no longer fit the 48 KB staging limit and are read from
    global memory; the grids are built for a 300 m x 40 m field."""
    rng = np.random.default_rng(5)
    obs = np.zeros((3000, 7))
    obs[:, 0] = rng.uniform(5, 300, 3000)
    obs[:, 1] = rng.choice([-1, 1], 3000) * rng.uniform(3.2, 20, 3000)
    obs[:, 2] = rng.uniform(0, np.pi, 3000)
    obs[:, 3] = rng.uniform(0.5, 3, 3000)
    obs[:, 4] = rng.uniform(0.5, 5, 3000)
    tree, par, smp, want = oracle_batch(obs, 7)
    got = run(clrrt, planner, obs, tree, par, smp)
    assert_rollouts_match(got, want, "3000 boxes")
    assert (want[:, 15] == 1).sum() > 20  # the scene does produce collisions


def test_pose_cell_overflow_falls_back(clrrt, planner):
    """More than 8 obstacles within reach of one pose cell (40 small boxes in a 2 m patch next to the lane): the lanes
    of those cells go through the position-grid lists."""
    rng = np.random.default_rng(9)
    obs = np.zeros((40, 7))
    obs[:, 0] = rng.uniform(14, 16, 40)
    obs[:, 1] = rng.uniform(1.6, 3.6, 40)
    obs[:, 2] = rng.uniform(0, np.pi, 40)
    obs[:, 3] = 0.4
    obs[:, 4] = 0.6
    obs = np.vstack([obs, scene_c1_boxes()])
    tree, par, smp, want = oracle_batch(obs, 3)
    got = run(clrrt, planner, obs, tree, par, smp)
    assert_rollouts_match(got, want, "dense patch")
    assert (want[:, 15] == 1).sum() > 20


def test_scene_far_from_the_origin(clrrt, planner):
    """Obstacle coordinates of several km (a car frame would never produce them, a world frame might): the broad phase
    works relative to the grid origin, subtracted in double."""
    off = np.array([4000.0, -7000.0])
    obs = scene_c1_boxes()
    obs[:, :2] += off
    car = (0, 0, 0, 0, 2, 0)
    goal = (50, 0, 0, 0)
    # the planner works in the car frame: move the whole problem instead — tree, samples and goal are shifted with the boxes
    tree, par, smp, want0 = oracle_batch(scene_c1_boxes(), 13, n=600)
    orc = CpuPlanner("oracle")
    orc.set_obstacles(obs)
    goal_s = (goal[0] + off[0], goal[1] + off[1], 0, 0)
    orc.tree_init(car, goal_s, 5.0)
    t2 = tree.copy()
    t2[:, 0] += off[0]; t2[:, 1] += off[1]; t2[:, 10] += off[0]; t2[:, 11] += off[1]; t2[:, 12] += off[0]; t2[:, 13] += off[1]
    orc.tree_import(t2)
    s2 = smp + off
    want = orc.rollout_batch(par, s2)
    got = run(clrrt, planner, obs, t2, par, s2, car, goal_s)
    assert_rollouts_match(got, want, "shifted scene")
    assert (want[:, 15] == 1).sum() > 5


def test_degenerate_boxes_and_many_moving(clrrt, planner):
    """Zero-size boxes, a box on top of the start pose, and 70 moving obstacles (more than one 64-entry chunk)."""
    rng = np.random.default_rng(2)
    mov = np.zeros((70, 7))
    mov[:, 0] = rng.uniform(10, 60, 70)
    mov[:, 1] = rng.choice([-1, 1], 70) * rng.uniform(2.5, 8, 70)
    mov[:, 2] = rng.uniform(0, np.pi, 70)
    mov[:, 3] = rng.uniform(0.5, 2, 70)
    mov[:, 4] = rng.uniform(0.5, 3, 70)
    mov[:, 5] = rng.uniform(-1.5, 0.5, 70)
    mov[:, 6] = rng.uniform(-0.3, 0.3, 70)
    zero = np.array([[20.0, 0.5, 0.3, 0.0, 0.0, 0.0, 0.0], [25.0, -0.5, 0.0, 0.0, 2.0, 0.0, 0.0]])
    obs = np.vstack([mov, zero, scene_c1_boxes()])
    tree, par, smp, want = oracle_batch(obs, 17)
    got = run(clrrt, planner, obs, tree, par, smp)
    assert_rollouts_match(got, want, "moving + degenerate")
    # a box over the car: every rollout from the root collides at its first step
    blocked = np.array([[1.4, 0.0, 0.0, 6.0, 12.0, 0.0, 0.0]])
    orc = CpuPlanner("oracle")
    orc.set_obstacles(blocked)
    orc.tree_init(CAR, GOAL, 5.0)
    s, h, _ = orc.draw_samples(64)
    want = orc.rollout_batch(np.zeros(64, np.int32), s)
    planner.set_query(CAR, GOAL, 5.0)
    planner.set_obstacles(blocked)
    planner.tree_reset(clrrt.root_node(CAR))
    got = clrrt.rollouts_as_table(planner.propagate_batch(np.zeros(64, np.int32), s))
    assert_rollouts_match(got, want, "blocked start")
    assert (got[:, 15] == 1).all() and (got[:, 14] == 1).all()


def test_non_finite_samples_and_full_tree(clrrt, planner):
    planner.set_query(CAR, GOAL, 5.0)
    planner.set_obstacles(scene_c1_boxes())
    planner.tree_reset(clrrt.root_node(CAR))
    s, h = clrrt.draw_samples(GOAL, 256, seed=4)
    s[5] = [np.nan, 1.0]
    s[9] = [np.inf, -np.inf]
    st = planner.expand_round(s, h)
    assert st.nodes_added > 0 and planner.tree_size() == 1 + st.nodes_added
    nodes = planner.tree_download()
    assert np.isfinite(nodes["state"][:, :7]).all()
    # a tree that cannot take all of the round's nodes: CLRRT_ERR_CAPACITY, and the prefix that fits IS appended — records
    # are in sample order, so these are the nodes the sequential reference would have added first (a query that fills
    # its tree keeps what it grew)
    s, h = clrrt.draw_samples(GOAL, 1024, seed=5)
    planner.tree_reset(clrrt.root_node(CAR))
    planner.expand_round(s, h)
    full = planner.tree_download()
    assert len(full) > 64
    small = clrrt.Planner(device=0, tree_capacity=64, max_round=1 << 10)
    small.set_query(CAR, GOAL, 5.0)
    small.set_obstacles(scene_c1_boxes())
    small.tree_reset(clrrt.root_node(CAR))
    with pytest.raises(clrrt.ClrrtError, match="capacity"):
        small.expand_round(s, h)
    assert small.tree_size() == 64
    assert small.tree_download().tobytes() == full[:64].tobytes()
    small.close()


def test_search_without_a_tree_and_with_non_finite_samples(clrrt):
    """A context that holds no tree yet refuses a search (CLRRT_ERR_STATE, whichever search is selected), and non-finite
    samples go through the sorted search like through the storage-order one: same lists, no candidates for them."""
    pl = clrrt.Planner(device=0, tree_capacity=1 << 12, max_round=1 << 10)
    try:
        pl.set_query(CAR, GOAL, 5.0)
        s, h = clrrt.draw_samples(GOAL, 64, seed=3)
        for mode in (0, 1, 2, 18):
            pl.set_nearest_mode(mode)
            with pytest.raises(clrrt.ClrrtError, match="-4"):
                pl.nearest_batch(s, h)
        pl.set_obstacles(scene_c1_boxes())
        pl.tree_reset(clrrt.root_node(CAR))
        pl.set_nearest_mode(0)
        for seed in (6, 7, 8):
            s2, h2 = clrrt.draw_samples(GOAL, 1024, seed=seed)
            pl.expand_round(s2, h2)
        s, h = clrrt.draw_samples(GOAL, 512, seed=9)
        s[3] = [np.nan, 1.0]; s[4] = [1.0, np.nan]; s[7] = [np.inf, 0.0]; s[8] = [-np.inf, np.inf]; s[11] = [1e30, -1e30]
        pl.set_nearest_mode(2)
        want = pl.nearest_batch(s, h)
        for mode in (1, 17, 18, 20):
            pl.set_nearest_mode(mode)
            got = pl.nearest_batch(s, h)
            for a, b in zip(got, want):
                assert np.array_equal(a, b), f"mode {mode}"
        assert (want[2][[3, 4]] == 0).all()   # NaN coordinates: nothing is feasible
    finally:
        pl.close()
