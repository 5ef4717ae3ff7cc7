"""The ROS-free C++ facade (cl-rrt_b200/host, libclrrt_host.so): MotionPlanner::planMotion end to end on the GPU,
checked against the whole-query golden replay of the reference (G3) at K=1."""
import ctypes as C
import os

import numpy as np
import pytest

import clrrt_b200 as clrrt
from cpulib import scene_c1_boxes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def plan(obstacles, samples_per_round, max_iterations, seed=1, car=(0, 0, 0, 0, 0, 0), goal=(50, 0, 0, 0)):
    lib = C.CDLL(os.path.join(ROOT, "cl-rrt_b200", "libclrrt_host.so"))
    obs = np.ascontiguousarray(obstacles, dtype=np.float64).reshape(-1, 7)
    tree, iters, tlen, blen = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    cnt = clrrt.Counters()
    traj = np.zeros((4096, 8))
    best = np.zeros(4096, np.int32)
    remat = C.c_double(0)
    vp, ip = C.c_void_p, C.c_int
    lib.clrrt_host_plan_motion.argtypes = [vp, vp, C.c_double, vp, ip, ip, ip, C.c_double, C.c_uint, ip, vp, vp, vp, vp, ip,
                                           vp, vp, ip, vp, vp]
    rc = lib.clrrt_host_plan_motion((C.c_double * 6)(*car), (C.c_double * 4)(*goal), 5.0,
                                    obs.ctypes.data if len(obs) else None, len(obs), samples_per_round, max_iterations,
                                    200.0, seed, 0, C.byref(tree), C.byref(iters), C.byref(cnt),
                                    traj.ctypes.data, 4096, C.byref(tlen), best.ctypes.data, 4096, C.byref(blen),
                                    C.byref(remat))
    assert rc == 0
    return dict(tree=tree.value, iters=iters.value, counters=cnt, traj=traj[:tlen.value], best=best[:blen.value],
                remat_err=remat.value)


@pytest.mark.parametrize("name", ["live", "obs"])
def test_plan_motion_k1_matches_reference_replay(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "g3_replay.npz"))
    r = plan(scene_c1_boxes() if name == "obs" else np.zeros((0, 7)), 1, 200)
    assert r["iters"] == 200 and r["tree"] == len(g[f"tree_{name}"])
    c = r["counters"]
    assert [c.fail_collision, c.fail_acclimit, c.fail_iterlimit, c.sim_count] == g[f"counters_{name}"].tolist()
    assert np.array_equal(r["best"], g[f"best_{name}"])
    # Node::tra re-materialised from (parent, sample, kind) ends exactly in the stored node state
    assert r["remat_err"] == 0.0
    # the published trajectory (generateMPCmessage + filterMPCmessage): waypoints >= 5 m apart along the best path,
    # starting at the car and ending within the goal tolerance of rrt/src/simulation.cpp:125
    t = r["traj"]
    assert len(t) >= 3
    d = np.hypot(np.diff(t[:, 0]), np.diff(t[:, 1]))
    # (the filter of motionplanner.cpp:130-150 emits the first two trajectory points, then one every >= 5 m)
    assert (d[1:] > 4.5).all() and t[0, 0] < 1.0
    end = g[f"tree_{name}"][g[f"best_{name}"][-1]]
    assert np.hypot(end[0] - 50, end[1]) <= 1.0


def test_plan_motion_rounds_and_time_budget():
    r = plan(scene_c1_boxes(), 2048, 3)
    assert r["tree"] > 500 and len(r["best"]) >= 2 and r["remat_err"] == 0.0
    r = plan(scene_c1_boxes(), 4096, -1)  # 200 ms wall-clock budget, as rrt/src/motionplanner.cpp:39
    assert r["iters"] >= 1 and r["tree"] > 1000
    print(f"200 ms query, 4096 samples/round: {r['iters']} rounds, {r['tree']} nodes, {r['counters'].sim_count} sim steps")


def test_simulation_with_the_reference_parameter_list():
    """clrrt::Simulation(RRT, state, ref, veh, GoalBiased, genProfile, Vstart) — the reference's own constructor signature
    (rrt/include/rrt/simulation.h:18-19): arbitrary 6-entry state, caller-owned curved reference, ref.v filled — equals the
    oracle's restatement of that constructor bit for bit; a failure reports through clrrt_host_last_error."""
    from cpulib import CpuPlanner
    lib = C.CDLL(os.path.join(ROOT, "cl-rrt_b200", "libclrrt_host.so"))
    lib.clrrt_host_last_error.restype = C.c_char_p
    vp, ip, dp = C.c_void_p, C.c_int, C.c_double
    lib.clrrt_host_simulate.argtypes = [vp, dp, vp, ip, ip, vp, ip, vp, vp, vp, ip, ip, ip, dp, vp, vp, vp]
    goal = np.array([50.0, 0, 0, 0])
    obs = scene_c1_boxes()
    orc = CpuPlanner("oracle")
    orc.set_obstacles(obs)
    orc.tree_init((0, 0, 0, 0, 2.5, 0), goal, 5.0)
    s = np.linspace(0, 45, 226)
    x, y = np.ascontiguousarray(s), np.ascontiguousarray(1.2 * np.sin(s / 7.0))
    for state, gb, gen in (([0.3, -0.4, 0.1, 0.02, 2.5, 0.3], 0, 1), ([0, 0.2, 0, 0, 1.0, 0], 1, 1), ([0, 0, 0.05, 0, 4.0, 0], 0, 0)):
        st = np.array(state, float)
        v = np.zeros(len(x)) if gen else np.full(len(x), 3.5)
        want, wv, wtraj = orc.simulate(np.concatenate([st, np.zeros(4)]), x, y, None if gen else v.copy(), gb=gb, gen_profile=gen, vstart=2.0)
        costs, flags, last = np.zeros(2), np.zeros(3, np.int32), np.zeros(10)
        n = lib.clrrt_host_simulate(goal.ctypes.data, 5.0, obs.ctypes.data, len(obs), 0, st.ctypes.data, 6, x.ctypes.data, y.ctypes.data,
                                    v.ctypes.data, len(x), gb, gen, 2.0, costs.ctypes.data, flags.ctypes.data, last.ctypes.data)
        assert n == len(wtraj), lib.clrrt_host_last_error()
        assert np.array_equal(last, want[:10]) and costs[0] == want[10] and costs[1] == want[11]
        assert flags.tolist() == [int(want[12]), int(want[13]), int(want[15])]
        assert np.array_equal(v, wv)
    # fewer than three reference points: an error code and a message, nothing thrown or printed
    rc = lib.clrrt_host_simulate(goal.ctypes.data, 5.0, None, 0, 0, st.ctypes.data, 6, x.ctypes.data, y.ctypes.data, v.ctypes.data, 2, 0, 1,
                                 2.0, None, None, None)
    assert rc < 0 and b"3 points" in lib.clrrt_host_last_error()
