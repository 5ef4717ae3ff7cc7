"""Live comparison of the C restatement with the reference's own sources (oracle/_ref, built from
/root/reference by oracle/build_ref.sh).  Skipped where the reference build is not present."""
import os

import numpy as np
import pytest

from cpulib import CpuPlanner, O_TAINT, ref_available, scene_c1_boxes

pytestmark = pytest.mark.skipif(not (ref_available() and ref_available(True)), reason="oracle/_ref not built")
NONE = np.zeros((0, 7))


def grow(kind, obs, seed, iters, car=(0, 0, 0, 0, 0, 0)):
    p = CpuPlanner(kind)
    if kind == "oracle":
        # Equal Dubins keys are common (one-step nodes share a pose) and the reference orders them with an
        # unstable std::sort; tie_mode 1 restates libstdc++'s algorithm so whole trees can be compared.
        p.set_tie_mode(1)
    p.set_obstacles(obs)
    p.srand(seed)
    p.tree_init(car)
    p.expand(iters)
    return p


@pytest.mark.parametrize("seed", [2, 7, 11])
@pytest.mark.parametrize("scene", ["live", "obs", "moving"])
def test_tree_growth_bit_exact_vs_defined_reference(seed, scene):
    obs = {"live": NONE, "obs": scene_c1_boxes(), "moving": scene_c1_boxes(moving=True)}[scene]
    a = grow("oracle", obs, seed, 300, (0, 0, 0, 0, 2, 0))
    b = grow("ref_defined", obs, seed, 300, (0, 0, 0, 0, 2, 0))
    assert np.array_equal(a.tree_export(), b.tree_export())
    assert a.counters() == b.counters()


def test_unmodified_reference_differs_only_when_tainted():
    obs = scene_c1_boxes()
    d = grow("ref_defined", obs, 3, 80)
    tree = d.tree_export()
    s, h, _ = d.draw_samples(600)
    cand, key, cnt = d.nearest_batch(s, h)
    ok = cnt > 0
    par, smp = cand[ok, 0], s[ok]
    out_d = d.rollout_batch(par, smp)
    u = CpuPlanner("ref")
    u.set_obstacles(obs)
    u.tree_init()
    u.tree_import(tree)
    out_u = u.rollout_batch(par, smp)
    o = CpuPlanner("oracle")
    o.set_obstacles(obs)
    o.tree_init()
    o.tree_import(tree)
    out_o = o.rollout_batch(par, smp)
    assert np.array_equal(out_o, out_d, equal_nan=True)
    differs = (out_u != out_d).any(axis=1)
    assert not (differs & (out_d[:, O_TAINT] == 0)).any()
    c2, k2, n2 = o.nearest_batch(s, h)
    assert np.array_equal(n2, cnt) and np.array_equal(k2, key)


def test_g5_golden_is_what_the_reference_produces(golden_dir):
    """The receding-horizon golden (config C5) replayed with the reference build present in this container."""
    import c5_scenario as sc
    if not ref_available(True):
        pytest.skip("oracle/_ref not built (no /root/reference)")
    g = np.load(os.path.join(golden_dir, "g5_replan.npz"))
    ref = CpuPlanner("ref_defined")
    ref.srand(1)
    ref.commit_reset()
    w = np.array([0, 0, 0, 0, 2.0, 0])
    t = 0.0
    for q in range(12):
        assert np.allclose(w, g["world"][q], rtol=0, atol=0)
        goal, obs = sc.to_car_frame(w, sc.world_goal(q), sc.world_obstacles(t))
        ref.set_obstacles(obs)
        carried, tree, nbest, steps, cost = ref.query_commit(w, goal, 5.0, int(g["iters"]))
        assert [carried, tree, nbest] == g["sizes"][q].tolist() and steps == int(g["sim_steps"][q])
        tr, rows = ref.best_traj()
        w = sc.advance(w, tr, rows[:nbest])
        t += sc.DT_QUERY
