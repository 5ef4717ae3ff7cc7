"""Config C5: the receding-horizon loop of MotionPlanner::planMotion with commit_path = true — tree initialised from the
previous best path (initializeTree, rrt/src/rrtplanner.cpp:50-94; transformNodesWorldToCar / CarToworld,
rrt/src/transformations.cpp:289-315), moving obstacles — through the C++ host facade on the GPU at K = 1, against the
golden loop G5 recorded from the reference's own sources (tests/golden/make_golden.py g5, tests/c5_scenario.py).
Every query gets the golden's inputs (world state, goal and obstacles in the car frame); the carried nodes are the
product's OWN previous best path, so a divergence anywhere shows up in every later query."""
import ctypes as C
import os

import numpy as np
import pytest

import clrrt_b200 as clrrt

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class HostPlanner:
    def __init__(self, samples_per_round=1, commit_path=True, device=0, tree_capacity=1 << 16):
        self.lib = C.CDLL(os.path.join(ROOT, "cl-rrt_b200", "libclrrt_host.so"))
        vp, ip, dp = C.c_void_p, C.c_int, C.c_double
        self.lib.clrrt_host_planner_create.restype = vp
        self.lib.clrrt_host_planner_create.argtypes = [ip, ip, ip, ip]
        self.lib.clrrt_host_planner_destroy.argtypes = [vp]
        self.lib.clrrt_host_planner_query.argtypes = [vp, vp, vp, dp, vp, ip, ip, dp, vp, vp, vp]
        self.lib.clrrt_host_planner_best_nodes.argtypes = [vp, vp, ip]
        self.lib.clrrt_host_planner_best_traj.argtypes = [vp, vp, ip, vp, ip]
        self.h = self.lib.clrrt_host_planner_create(device, samples_per_round, int(commit_path), tree_capacity)

    def query(self, world6, goal4, obstacles, max_iterations, vmax=5.0, budget_ms=200.0):
        w = np.ascontiguousarray(world6, dtype=np.float64)
        g = np.ascontiguousarray(goal4, dtype=np.float64)
        o = np.ascontiguousarray(obstacles, dtype=np.float64).reshape(-1, 7)
        sizes = np.zeros(4, np.int32)
        cost = C.c_double(0)
        cnt = clrrt.Counters()
        rc = self.lib.clrrt_host_planner_query(self.h, w.ctypes.data, g.ctypes.data, vmax, o.ctypes.data if len(o) else None,
                                               len(o), max_iterations, budget_ms, sizes.ctypes.data, C.byref(cost), C.byref(cnt))
        assert rc == 0
        return sizes, cost.value, cnt

    def best_nodes(self, cap=64):
        rec = np.zeros((cap, 20))
        n = self.lib.clrrt_host_planner_best_nodes(self.h, rec.ctypes.data, cap)
        return rec[:min(n, cap)]

    def best_traj(self, cap_rows=40000, cap_nodes=64):
        tr = np.zeros((cap_rows, 10))
        rows = np.zeros(cap_nodes, np.int32)
        n = self.lib.clrrt_host_planner_best_traj(self.h, tr.ctypes.data, cap_rows, rows.ctypes.data, cap_nodes)
        return tr[:min(n, cap_rows)], rows

    def close(self):
        if self.h:
            self.lib.clrrt_host_planner_destroy(self.h)
            self.h = None


# Every one of the 100 recorded queries has to equal the reference's.  Two things this loop exercises that nothing else
# does: query 41 holds a candidate list with two exactly equal keys (std::sort's order decides which parent is tried first),
# and in query 50 a goal-biased rollout crosses the duplicated junction point of its reference with the controller's
# 3-point window straddling the pair — the Lagrange interpolation divides by ~1e-15 and the saturated steer command takes a
# sign that hangs on the last bit of sin/cos/tan, which is why the kernels restate glibc's double
# routines (csrc/refmath64.cuh).  With CUDA's own libm that rollout ends two steps later and the loops part ways there.
FIRST_JUNCTION_NOISE_QUERY = 100  # = no query is exempt


def test_receding_horizon_loop_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "g5_replan.npz"))
    iters = int(g["iters"])
    n_queries = int(g["world"].shape[0])  # config C5: all 100 recorded queries (~2 ms per K=1 expandTree call on the GPU)
    hp = HostPlanner(samples_per_round=1, commit_path=True)
    C.CDLL(None).srand(C.c_uint(1))  # the reference never seeds rand(); its stream continues across queries
    worst = 0.0
    exact = 0
    for q in range(n_queries):
        sizes, cost, cnt = hp.query(g["world"][q], g["goal"][q], g["obstacles"][q], iters)
        want = g["sizes"][q]
        nodes = hp.best_nodes()
        ref_nodes = g["nodes"][q][:int(g["n_nodes"][q])]
        same = (sizes[:3].tolist() == want.tolist() and cnt.sim_count == int(g["sim_steps"][q]) and len(nodes) == len(ref_nodes)
                and np.array_equal(nodes[:, [7, 17, 18, 19]], ref_nodes[:, [7, 17, 18, 19]]))
        if not same and q >= FIRST_JUNCTION_NOISE_QUERY:
            # past the noise event the two loops are different (equally valid) realisations: the planner still has to
            # find a path in every query and stay close to the reference's amount of work
            assert sizes[2] > 0 and sizes[3] == iters, f"query {q}: no path"
            assert abs(cnt.sim_count - int(g["sim_steps"][q])) <= 0.5 * int(g["sim_steps"][q]), f"query {q}: sim steps {cnt.sim_count}"
            continue
        assert sizes[:3].tolist() == want.tolist(), f"query {q}: carried/tree/best {sizes[:3]} vs reference {want}"
        assert sizes[3] == iters and cnt.sim_count == int(g["sim_steps"][q]), f"query {q}: sim steps"
        assert len(nodes) == len(ref_nodes)
        # discrete fields exactly (parent, goal flag, reference length, waypoint index), the rest to 1e-6 relative
        assert np.array_equal(nodes[:, [7, 17, 18, 19]], ref_nodes[:, [7, 17, 18, 19]]), f"query {q}: node bookkeeping"
        cont = [0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16]
        err = np.abs(nodes[:, cont] - ref_nodes[:, cont]) / np.maximum(1.0, np.abs(ref_nodes[:, cont]))
        if q < FIRST_JUNCTION_NOISE_QUERY:
            worst = max(worst, float(err.max()))
            assert err.max() < 1e-6, f"query {q}: node values differ by {err.max()}"
            assert abs(cost - float(g["cost"][q])) <= 1e-6 * max(1.0, abs(float(g["cost"][q])))
            tr, rows = hp.best_traj()
            assert rows[:len(nodes)].tolist() == g["rows"][q][:len(nodes)].tolist(), f"query {q}: trajectory lengths"
            exact += 1
    assert exact == FIRST_JUNCTION_NOISE_QUERY
    print(f"C5: the first {exact} consecutive queries equal to the reference (max relative deviation {worst:.2e}), including the "
          f"query whose candidate list depends on std::sort's order of equal keys; {n_queries} queries run")
    hp.close()


def test_single_sample_search_orders_equal_keys_like_std_sort(golden_dir):
    """Query 41 of G5 holds an iteration whose two best parents have exactly equal keys; the reference tries them in the
    order its std::sort leaves (the higher node id first there).  Up to that query nothing else distinguishes the tie
    rules, so sim steps of query 41 separate them: 5128 (reference, tie mode 1) against 5135 (lower node id first)."""
    g = np.load(os.path.join(golden_dir, "g5_replan.npz"))
    hp = HostPlanner(samples_per_round=1, commit_path=True)
    C.CDLL(None).srand(C.c_uint(1))
    for q in range(42):
        sizes, cost, cnt = hp.query(g["world"][q], g["goal"][q], g["obstacles"][q], int(g["iters"]))
    assert cnt.sim_count == int(g["sim_steps"][41]) == 5128
    hp.close()


def test_receding_horizon_rounds_keep_a_path():
    """The same loop with 2048 samples per round (the batched formulation): every query finds a path, the tree is
    initialised from the previous one, and the car makes progress."""
    import c5_scenario as sc
    hp = HostPlanner(samples_per_round=2048, commit_path=True, tree_capacity=1 << 18)
    C.CDLL(None).srand(C.c_uint(1))
    w = np.array([0, 0, 0, 0, 2.0, 0])
    t = 0.0
    carried_any = False
    for q in range(25):
        goal, obs = sc.to_car_frame(w, sc.world_goal(q), sc.world_obstacles(t))
        sizes, cost, cnt = hp.query(w, goal, obs, 2)
        assert sizes[2] >= 2 and sizes[1] > 100 and sizes[0] >= 1, f"query {q}: {sizes}"
        carried_any = carried_any or sizes[0] > 1
        tr, rows = hp.best_traj()
        w = sc.advance(w, tr, rows[:sizes[2]])
        t += sc.DT_QUERY
    assert w[0] > 15.0 and carried_any
    hp.close()
