"""GPU parity of whole expansion rounds (clrrt_expand_round): K=1 reproduces the reference's sequential
expandTree; K>1 rounds against one snapshot match the oracle's statement of the same batched semantics."""
import os

import numpy as np
import pytest

from cpulib import CpuPlanner, scene_c1_boxes
from gpu_common import clrrt, rel_err  # noqa: F401

pytestmark = pytest.mark.gpu
NONE = np.zeros((0, 7))
DISC_NODE = [7, 17, 18, 19]  # IDwp, parent, goal flag, n_ref


@pytest.fixture(scope="module")
def planner(clrrt):
    pl = clrrt.Planner(device=0, tree_capacity=1 << 16, max_round=1 << 13)
    yield pl
    pl.close()


def gb_descendants(tree):
    """Nodes produced by a goal-biased rollout (two-segment reference) and everything below them: their states
    carry the junction-point rounding noise described in tests/test_gpu_rollout.py."""
    n = len(tree)
    mark = np.zeros(n, bool)
    for i in range(1, n):
        p = int(tree[i, 17])
        # a goal-biased child starts where its parent's reference ended and ends 4.2 m beyond the goal alignment point
        is_gb = i > 0 and p == i - 1 and tree[i, 10] == tree[p, 12] and tree[i, 11] == tree[p, 13] and abs(tree[i, 12] - 53.2) < 1e-6
        mark[i] = is_gb or mark[p]
    return mark


@pytest.mark.parametrize("name", ["live", "obs"])
def test_g3_whole_query_replay_k1(clrrt, planner, golden_dir, name):
    g = np.load(os.path.join(golden_dir, "g3_replay.npz"))
    planner.set_query((0, 0, 0, 0, 0, 0), (50, 0, 0, 0), 5.0)
    planner.set_obstacles(scene_c1_boxes() if name == "obs" else NONE)
    planner.tree_reset(clrrt.root_node((0, 0, 0, 0, 0, 0)))
    s, h = clrrt.draw_samples((50, 0, 0, 0), 200, seed=1)
    assert np.array_equal(s, g["samples"]) and np.array_equal(h, g["heuristic"])
    for j in range(200):
        planner.expand_round(s[j:j + 1], h[j:j + 1])
    got, want = planner.tree_download_records(), g[f"tree_{name}"]
    assert len(got) == len(want)
    assert np.array_equal(got[:, DISC_NODE], want[:, DISC_NODE])
    noisy = gb_descendants(want)
    assert rel_err(got[~noisy], want[~noisy]).max() < 1e-6
    assert np.abs(got[noisy][:, :3] - want[noisy][:, :3]).max() < 0.05
    # since the device evaluates glibc's double sin/cos/tan (csrc/refmath64.cuh) even the goal-biased nodes are bit-equal
    assert rel_err(got, want).max() == 0.0
    c = planner.counters()
    assert [c["fail_collision"], c["fail_acclimit"], c["fail_iterlimit"], c["sim_count"]] == g[f"counters_{name}"].tolist()
    assert np.array_equal(planner.best_path(), g[f"best_{name}"])
    print(f"{name}: {len(got)} nodes, {int(noisy.sum())} goal-biased/descendant nodes compared loosely")


@pytest.mark.parametrize("K", [1, 7, 256, 4096])
def test_snapshot_rounds_vs_oracle(clrrt, planner, K):
    car, goal = (0, 0, 0, 0, 2, 0), (50, 0, 0, 0)
    orc = CpuPlanner("oracle")
    orc.set_obstacles(scene_c1_boxes())
    orc.srand(5)
    orc.tree_init(car, goal, 5.0)
    orc.expand(50)
    planner.set_query(car, goal, 5.0)
    planner.set_obstacles(scene_c1_boxes())
    planner.tree_reset_records(orc.tree_export())
    for r in range(3):
        s, h, _ = orc.draw_samples(K)
        orc.expand_round(s, h)
        st = planner.expand_round(s, h)
        a, b = planner.tree_download_records(), orc.tree_export()
        assert len(a) == len(b) == st.tree_size
        assert np.array_equal(a[:, DISC_NODE], b[:, DISC_NODE])
        noisy = gb_descendants(b)
        assert rel_err(a[~noisy], b[~noisy]).max() < 1e-6
        assert rel_err(a, b).max() == 0.0  # bit-equal, goal-biased nodes included (csrc/refmath64.cuh)
    oc = orc.counters()
    gc = planner.counters()
    # counters restart at tree_reset on the GPU side; the oracle's include the 50 set-up iterations
    assert gc["sim_count"] > 0 and gc["fail_collision"] <= oc["fail_collision"]


def test_empty_and_degenerate_rounds(clrrt, planner):
    car, goal = (0, 0, 0, 0, 0, 0), (50, 0, 0, 0)
    planner.set_query(car, goal, 5.0)
    planner.set_obstacles(NONE)
    planner.tree_reset(clrrt.root_node(car))
    # samples behind the car: no feasible parent -> nothing appended, no rollout run
    st = planner.expand_round([[-5.0, 0.0], [-3.0, 1.0]], [0, 1])
    assert st.nodes_added == 0 and st.rollouts == 0 and planner.tree_size() == 1
    with pytest.raises(clrrt.ClrrtError):
        planner.expand_round(np.zeros((planner.max_round + 1, 2)), np.zeros(planner.max_round + 1, np.uint8))


def test_async_download_equals_sync_download(clrrt, planner):
    """clrrt_tree_download_range_async: the nodes staged before the range is truncated and re-grown by the next round arrive
    unchanged (the copy runs on a second stream beside that round)."""
    import torch
    car, goal = (0, 0, 0, 0, 2, 0), (50, 0, 0, 0)
    planner.set_query(car, goal, 5.0)
    planner.set_obstacles(scene_c1_boxes())
    planner.tree_reset(clrrt.root_node(car))
    s, h = clrrt.draw_samples(goal, 3 * 4096, seed=9)
    planner.expand_round(s[:4096], h[:4096])
    n0 = planner.tree_size()
    pinned = [torch.empty(2 * 4096 * clrrt.RECORD_BYTES, dtype=torch.uint8).pin_memory().numpy().view(clrrt.NODE_DTYPE) for _ in range(2)]
    want = []
    for r in range(1, 3):
        planner.expand_round(s[r * 4096:(r + 1) * 4096], h[r * 4096:(r + 1) * 4096])
        n = planner.tree_size() - n0
        want.append(planner.tree_download_range(n0, n).copy())
        got = planner.tree_download_range_async(n0, n, pinned[r - 1])
        planner.tree_truncate(n0)                                            # the next round overwrites the range at once
        if r == 1:
            continue
    planner.download_wait()
    assert pinned[1][:len(want[1])].tobytes() == want[1].tobytes()
    assert pinned[0][:len(want[0])].tobytes() == want[0].tobytes()
