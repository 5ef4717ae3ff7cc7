"""GPU parity of the candidate search (clrrt_nearest_batch) with sortNodesExplore / sortNodesOptimize."""
import os

import numpy as np
import pytest

from cpulib import CpuPlanner, check_candidate_lists, scene_c1_boxes
from gpu_common import clrrt  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def planner(clrrt):
    pl = clrrt.Planner(device=0, tree_capacity=1 << 14, max_round=1 << 13)
    yield pl
    pl.close()


def ulp_diff(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


# storage-order search / nodes and samples sorted along the goal bearing first / sorted into (axis slab, lateral bin) cells
# with 4 and 16 lateral bins (the order large trees get; forced here on small ones)
@pytest.mark.parametrize("mode", [2, 1, 18, 20])
@pytest.mark.parametrize("N", [1, 64, 250, 1500])
def test_g2_candidate_lists(clrrt, planner, golden_dir, N, mode):
    g = np.load(os.path.join(golden_dir, "g2_nearest.npz"))
    car, goal = (0, 0, 0, 0, 3, 0), (50, 0, 0, 0)
    planner.set_query(car, goal, 5.0)
    planner.tree_reset_records(g["tree"][:N])
    planner.set_nearest_mode(mode)
    try:
        cand, key, cnt = planner.nearest_batch(g["samples"], g["heuristic"])
    finally:
        planner.set_nearest_mode(0)
    assert np.array_equal(cnt, g[f"count_{N}"])
    # keys: bit-exact with the reference's float libm
    assert np.array_equal(key, g[f"key_{N}"]), f"max ulp diff {ulp_diff(key, g[f'key_{N}']).max()}"
    orc = CpuPlanner("oracle")
    orc.tree_init(car, goal, 5.0)
    orc.tree_import(g["tree"][:N])
    check_candidate_lists(orc, g["samples"], g["heuristic"], cand, key, cnt)
    # where no key tie is involved the node ids are the reference's
    untied = np.ones_like(cand, bool)
    untied[:, 1:] &= key[:, 1:] != key[:, :-1]
    untied[:, :-1] &= key[:, :-1] != key[:, 1:]
    untied[:, -1] = False
    assert np.array_equal(cand[untied], g[f"cand_{N}"][untied])


def test_large_snapshot_vs_oracle(clrrt, planner):
    """4096 samples x 4096 nodes (config sizes) against the oracle; ties resolved towards the lower node id on
    both sides, so the lists must be identical."""
    car, goal = (0, 0, 0, 0, 2, 0), (50, 0, 0, 0)
    orc = CpuPlanner("oracle")
    orc.set_obstacles(scene_c1_boxes())
    orc.srand(3)
    orc.tree_init(car, goal, 5.0)
    orc.expand(100)
    s, h, _ = orc.draw_samples(4096)
    planner.set_query(car, goal, 5.0)
    planner.set_obstacles(scene_c1_boxes())
    planner.tree_reset_records(orc.tree_export())
    while planner.tree_size() < 4096:  # grow on the GPU, then hand the same tree to the oracle
        s2, h2, _ = orc.draw_samples(2048)
        planner.expand_round(s2, h2)
    planner.tree_truncate(4096)
    tree = planner.tree_download_records()
    orc.tree_import(tree)
    oc, ok, on = orc.nearest_batch(s, h)
    for mode in (2, 1, 0, 17, 18, 20, 21):
        planner.set_nearest_mode(mode)
        cand, key, cnt = planner.nearest_batch(s, h)
        assert np.array_equal(cnt, on)
        assert np.array_equal(key, ok)
        assert np.array_equal(cand, oc)
    # a goal off the x axis: the sort axis follows the goal bearing
    goal2 = (40, 25, 0.5, 0)
    planner.set_query(car, goal2, 5.0)
    planner.tree_reset_records(tree)
    orc.tree_init(car, goal2, 5.0)
    orc.tree_import(tree)
    oc, ok, on = orc.nearest_batch(s, h)
    for mode in (2, 1, 18, 20):
        planner.set_nearest_mode(mode)
        cand, key, cnt = planner.nearest_batch(s, h)
        assert np.array_equal(cnt, on) and np.array_equal(key, ok) and np.array_equal(cand, oc)
    planner.set_nearest_mode(0)


def test_spatial_orders_agree_on_a_large_tree(clrrt, golden_dir):
    """A tree of ~10^5 nodes (the 200 ms query of config C1): the storage-order search (every tile, every node) against the
    sorted searches with 1, 4 and 16 lateral bins and the default.  With the optimise key the best parents of many samples
    are the root and its copies, tens of metres behind the sample: a search that closes a direction on anything but the
    axis distance loses them."""
    import bench
    K = 16384
    pl = clrrt.Planner(device=0, tree_capacity=(1 << 18), max_round=K)
    try:
        pl.set_query(bench.C1_CAR, bench.C1_GOAL, bench.VMAX)
        pl.set_obstacles(bench.scene_c1_boxes())
        pl.tree_reset(clrrt.root_node(bench.C1_CAR))
        clrrt.draw_samples(bench.C1_GOAL, 1, seed=1)
        while pl.tree_size() < 100000:
            s, h = clrrt.draw_samples(bench.C1_GOAL, K)
            pl.expand_round(s, h)
        s, h = clrrt.draw_samples(bench.C1_GOAL, 4096)
        assert h.any() and not h.all()
        pl.set_nearest_mode(2)
        want = pl.nearest_batch(s, h)
        for mode in (16, 18, 20, 0):
            pl.set_nearest_mode(mode)
            got = pl.nearest_batch(s, h)
            for a, b in zip(got, want):
                assert np.array_equal(a, b), f"mode {mode}"
    finally:
        pl.close()


def test_equal_keys_follow_std_sort_for_single_sample_searches(clrrt, planner, golden_dir):
    """A tree with exact copies of some nodes (equal keys by construction, as carried-over trees produce them): searched
    one sample at a time, the candidate lists must equal the oracle's in BOTH tie rules — tie_mode 1 (default): the order
    libstdc++'s std::sort leaves (the oracle restates its introsort, pinned against the reference binary); tie_mode 0:
    lower node id first.  Batched searches always use the node-id rule."""
    g = np.load(os.path.join(golden_dir, "g2_nearest.npz"))
    car, goal = (0, 0, 0, 0, 3, 0), (50, 0, 0, 0)
    base = g["tree"][:120]
    dup = base[[7, 7, 30, 55, 55, 55, 90]].copy()      # copies of nodes, appended with new ids
    tree = np.concatenate([base, dup])
    smp, heu = g["samples"][:200], g["heuristic"][:200]
    planner.set_query(car, goal, 5.0)
    planner.tree_reset_records(tree)
    orc = CpuPlanner("oracle")
    orc.tree_init(car, goal, 5.0)
    orc.tree_import(tree)
    differ = 0
    try:
        lists = {}
        for mode in (1, 0):
            planner.set_tie_mode(mode)
            orc.set_tie_mode(mode)
            oc, ok, on = orc.nearest_batch(smp, heu)
            got = [planner.nearest_batch(smp[j:j + 1], heu[j:j + 1]) for j in range(len(smp))]
            gc = np.concatenate([x[0] for x in got]); gk = np.concatenate([x[1] for x in got]); gn = np.concatenate([x[2] for x in got])
            assert np.array_equal(gn, on) and np.array_equal(gk, ok)
            assert np.array_equal(gc, oc), f"tie mode {mode}: candidate order differs from the oracle's"
            lists[mode] = gc
        differ = int((lists[0] != lists[1]).any(1).sum())
        # batched search: node-id rule whatever the mode
        planner.set_tie_mode(1)
        bc, bk, bn = planner.nearest_batch(smp, heu)
        assert np.array_equal(bc, lists[0])
    finally:
        planner.set_tie_mode(1)
    assert planner.tie_sorts() > 0
    print(f"equal keys: {differ} of {len(smp)} lists differ between the two tie rules; {planner.tie_sorts()} searches repeated std::sort on the host")


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1])
def test_sorted_search_bounds_on_an_adversarial_tree(clrrt, seed):
    """The tile bounds of the sorted search (box, projected cost, feasibility by direction class) only ever skip what the
    exact tests reject — on a tree that breaks every regularity a grown tree has: references in all directions (also of zero
    length, also ending far from their node), costs unrelated to the distance from the root (zero, huge, negative), nodes
    outside the binned range, duplicated nodes.  Storage-order search (every node of every tile) = sorted search, for both
    keys and 1, 4, 16 lateral bins."""
    import bench
    rng = np.random.default_rng(seed)
    n, K = 120000, 2048
    nodes = np.zeros(n, clrrt.NODE_DTYPE)
    nodes[0] = clrrt.root_node(bench.C1_CAR)[0]
    x = rng.uniform(-20.0, 80.0, n); y = rng.uniform(-12.0, 12.0, n)
    x[:50] = rng.uniform(-300.0, 300.0, 50)            # outside the bins
    th = rng.uniform(-np.pi, np.pi, n)
    st = nodes["state"]; st[1:, 0] = x[1:]; st[1:, 1] = y[1:]; st[1:, 2] = th[1:]; st[1:, 4] = 5.0
    ang = rng.uniform(-np.pi, np.pi, n)                # direction of the node's own reference: anything
    ang[rng.random(n) < 0.7] *= 0.2                    # most roughly forward, like a grown tree
    lead = rng.uniform(0.0, 8.0, n); length = rng.uniform(0.0, 6.0, n)
    length[rng.random(n) < 0.01] = 0.0                 # zero-length references
    rb = np.stack([x + lead * np.cos(th), y + lead * np.sin(th)], 1)
    rf = rb - length[:, None] * np.stack([np.cos(ang), np.sin(ang)], 1)
    nodes["ref_back"][1:] = rb[1:]; nodes["ref_front"][1:] = rf[1:]; nodes["ref_vback"][1:] = 5.0
    ce = np.hypot(x, y) + rng.exponential(0.5, n)
    ce[rng.random(n) < 0.05] = 0.0
    ce[rng.random(n) < 0.02] = 1.0e6
    ce[rng.random(n) < 0.02] *= -1.0
    nodes["costE"][1:] = ce[1:].astype(np.float32)
    nodes["parent"][1:] = 0
    nodes["n_ref"][1:] = 2
    nodes[n // 2:n // 2 + 500] = nodes[1000:1500]      # duplicates: equal keys, the lower node id wins
    nodes["parent"][n // 2:n // 2 + 500] = 0
    pl = clrrt.Planner(device=0, tree_capacity=n + 16, max_round=K)
    try:
        pl.set_query(bench.C1_CAR, bench.C1_GOAL, bench.VMAX)
        pl.tree_reset(nodes)
        s = np.stack([rng.uniform(-5.0, 70.0, K), rng.uniform(-9.0, 9.0, K)], 1)
        for h in (np.zeros(K, np.uint8), np.ones(K, np.uint8), (rng.random(K) < 0.3).astype(np.uint8)):
            pl.set_nearest_mode(2)
            want = pl.nearest_batch(s, h)
            assert (want[2] > 0).mean() > 0.5
            for mode in (16, 18, 20, 0):
                pl.set_nearest_mode(mode)
                got = pl.nearest_batch(s, h)
                for a, b in zip(got, want):
                    assert np.array_equal(a, b), f"mode {mode}"
    finally:
        pl.close()
