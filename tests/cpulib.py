"""ctypes loader shared by the tests, bench.py's CPU legs and the golden generator.

Wraps the two CPU implementations that expose the same flat C interface:
  * prefix "orc": oracle/libclrrt_oracle.so   (plain-C restatement, oracle/clrrt_oracle.c)
  * prefix "ref": oracle/_ref/libclrrt_ref*.so (the reference's own sources, oracle/ref_driver.cpp)
TEST INFRASTRUCTURE: nothing under cl-rrt_b200/ imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
NODE_STRIDE = 20
OUT_STRIDE = 24
# columns of a rollout record (ref_driver.cpp fill_out / clrrt_oracle.c sim_to_out)
O_STATE, O_COSTE, O_COSTS, O_END, O_GOAL, O_NSTEPS, O_FAIL, O_NREF = slice(0, 10), 10, 11, 12, 13, 14, 15, 16
O_REFBX, O_REFBY, O_VBACK, O_TAINT, O_TRACE, O_IDWP0 = 17, 18, 19, 20, 21, 22

_dp = C.POINTER(C.c_double)


def _ptr(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libclrrt_oracle.so"])


def ref_available(defined=False):
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libclrrt_ref_defined.so" if defined else "libclrrt_ref.so"))


class CpuPlanner:
    """One CPU implementation.  The libraries keep file-scope state (like the reference): one planner per process
    per library."""

    def __init__(self, kind="oracle"):
        if kind == "oracle":
            path, self.p = os.path.join(ORACLE_DIR, "libclrrt_oracle.so"), "orc"
            if not os.path.exists(path):
                build_oracle()
        elif kind == "ref":
            path, self.p = os.path.join(ORACLE_DIR, "_ref", "libclrrt_ref.so"), "ref"
        elif kind == "ref_defined":
            path, self.p = os.path.join(ORACLE_DIR, "_ref", "libclrrt_ref_defined.so"), "ref"
        else:
            raise ValueError(kind)
        self.kind = kind
        self.lib = C.CDLL(path)
        f = self._f
        f("rollout_batch").restype = C.c_double
        f("nearest_batch").restype = C.c_double
        f("obb_dist").restype = C.c_double
        f("obs_distance").restype = C.c_double
        f("get_ref_res").restype = C.c_double
        f("dubins").restype = C.c_float
        f("dubins").argtypes = [C.c_double] * 5 + [C.c_int]
        f("tree_init").argtypes = [C.c_void_p, C.c_void_p, C.c_double]
        f("expand_timed").argtypes = [C.c_double, C.c_void_p]
        f("init")()

    def _f(self, name):
        return getattr(self.lib, f"{self.p}_{name}")

    # ---- configuration -------------------------------------------------------------------------------
    def srand(self, seed):
        self._f("srand")(C.c_uint(seed))

    def set_obstacles(self, obs):
        obs = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1, 7)
        self._f("set_obstacles")(_ptr(obs), C.c_int(len(obs)))

    def set_road(self, bend, Cxy3=(0, 0, 0), lane_shift=0.0):
        """MotionRequest.bend / Cxy / laneShifts[0]: the lane-deviation cost of the curved-road mode (takes effect at
        the next tree_init for the reference libraries)."""
        c = np.ascontiguousarray(Cxy3, dtype=np.float64)
        self._f("set_road")(C.c_int(int(bend)), _ptr(c), C.c_double(lane_shift))

    def set_tie_mode(self, mode):
        """oracle only: 0 = equal keys ordered by node id (the product's rule for batched searches), 1 = libstdc++ std::sort order
        (the reference's, and the product's for single-sample searches)."""
        self._f("set_tie_mode")(C.c_int(mode))

    def set_weights(self, w5):
        w = np.ascontiguousarray(w5, dtype=np.float64)
        self._f("set_weights")(_ptr(w))

    def vehicle(self):
        v = np.zeros(14)
        self._f("get_vehicle")(_ptr(v))
        return v

    def tree_init(self, car_state6=(0, 0, 0, 0, 0, 0), goal=(50, 0, 0, 0), vmax=5.0):
        cs = np.ascontiguousarray(car_state6, dtype=np.float64)
        g = np.ascontiguousarray(goal, dtype=np.float64)
        self._f("tree_init")(_ptr(cs), _ptr(g), C.c_double(vmax))

    def ref_res(self):
        return self._f("get_ref_res")()

    # ---- sequential reference algorithm ---------------------------------------------------------------
    def expand(self, iters):
        return self._f("expand")(C.c_int(iters))

    def expand_timed(self, budget_ms):
        it = C.c_int(0)
        n = self._f("expand_timed")(C.c_double(budget_ms), C.byref(it))
        return n, it.value

    # ---- receding-horizon queries with a carried tree (reference libraries only, config C5) --------------
    def commit_reset(self):
        self._f("commit_reset")()

    def query_commit(self, world_state6, goal4_car, vmax, iters):
        """One planMotion query with commit_path = true (see oracle/ref_driver.cpp: ref_query_commit).
        Returns (carried nodes in the initial tree, final tree size, best path nodes, sim steps, best cost)."""
        w = np.ascontiguousarray(world_state6, dtype=np.float64)
        g = np.ascontiguousarray(goal4_car, dtype=np.float64)
        sizes = (C.c_int * 4)()
        cost = C.c_double(0)
        self._f("query_commit")(_ptr(w), _ptr(g), C.c_double(vmax), C.c_int(iters), sizes, C.byref(cost))
        return sizes[0], sizes[1], sizes[2], sizes[3], cost.value

    def best_nodes(self, cap=256):
        rec = np.zeros((cap, NODE_STRIDE))
        n = self._f("best_nodes")(_ptr(rec), C.c_int(cap))
        return rec[:min(n, cap)]

    def best_traj(self, cap_rows=20000, cap_nodes=256):
        tr = np.zeros((cap_rows, 10))
        rows = np.zeros(cap_nodes, np.int32)
        n = self._f("best_traj")(_ptr(tr), C.c_int(cap_rows), _ptr(rows), C.c_int(cap_nodes))
        return tr[:min(n, cap_rows)], rows

    def expand_with(self, samples, heuristic):
        """oracle only: expandTree per sample, sequentially (K=1 semantics), with caller-supplied draws."""
        s = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, 2)
        h = np.ascontiguousarray(heuristic, dtype=np.uint8)
        return self._f("expand_with")(_ptr(s), _ptr(h), C.c_int(len(s)))

    def expand_round(self, samples, heuristic):
        """oracle only: all samples against one tree snapshot, appended in sample order."""
        s = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, 2)
        h = np.ascontiguousarray(heuristic, dtype=np.uint8)
        return self._f("expand_round")(_ptr(s), _ptr(h), C.c_int(len(s)))

    def tree_size(self):
        return self._f("tree_size")()

    def counters(self):
        c = (C.c_int * 4)()
        self._f("counters")(c)
        return dict(fail_collision=c[0], fail_acclimit=c[1], fail_iterlimit=c[2], sim_count=c[3])

    def tree_export(self):
        n = self.tree_size()
        out = np.zeros((n, NODE_STRIDE))
        self._f("tree_export")(_ptr(out), C.c_int(n))
        return out

    def tree_import(self, nodes):
        nodes = np.ascontiguousarray(nodes, dtype=np.float64).reshape(-1, NODE_STRIDE)
        self._f("tree_import")(_ptr(nodes), C.c_int(len(nodes)))

    def draw_samples(self, K):
        s = np.zeros((K, 2))
        h = np.zeros(K, dtype=np.uint8)
        r = np.zeros(K)
        self._f("draw_samples")(C.c_int(K), _ptr(s), _ptr(h), _ptr(r))
        return s, h, r

    # ---- batched primitives ---------------------------------------------------------------------------
    def rollout_batch(self, parent, samples, gb=None):
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        samples = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, 2)
        M = len(parent)
        gbp = None if gb is None else np.ascontiguousarray(gb, dtype=np.uint8)
        out = np.zeros((M, OUT_STRIDE))
        secs = self._f("rollout_batch")(_ptr(parent), _ptr(samples), None if gbp is None else _ptr(gbp), C.c_int(M),
                                        _ptr(out))
        self.last_seconds = secs
        return out

    def rollout_traj(self, parent, sample, gb=0, cap=512):
        s = np.ascontiguousarray(sample, dtype=np.float64)
        traj = np.zeros((cap, 10))
        refv = np.zeros(2048)
        n = self._f("rollout_traj")(C.c_int(parent), _ptr(s), C.c_int(gb), _ptr(traj), C.c_int(cap), _ptr(refv),
                                    C.c_int(2048))
        return traj[:n], refv

    def nearest_batch(self, samples, heuristic):
        samples = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, 2)
        heuristic = np.ascontiguousarray(heuristic, dtype=np.uint8)
        K = len(samples)
        cand = np.zeros((K, 10), dtype=np.int32)
        key = np.zeros((K, 10), dtype=np.float32)
        cnt = np.zeros(K, dtype=np.int32)
        secs = self._f("nearest_batch")(_ptr(samples), _ptr(heuristic), C.c_int(K), _ptr(cand), _ptr(key), _ptr(cnt))
        self.last_seconds = secs
        return cand, key, cnt

    def keys(self, sample, heuristic):
        n = self.tree_size()
        s = np.ascontiguousarray(sample, dtype=np.float64)
        key = np.zeros(n, dtype=np.float32)
        feas = np.zeros(n, dtype=np.uint8)
        self._f("keys")(_ptr(s), C.c_int(int(heuristic)), _ptr(key), _ptr(feas))
        return key, feas

    def dubins(self, sx, sy, nx, ny, nth, direction=1):
        return self._f("dubins")(sx, sy, nx, ny, nth, direction)

    def obb_dist(self, a5, b5):
        a = np.ascontiguousarray(a5, dtype=np.float64)
        b = np.ascontiguousarray(b5, dtype=np.float64)
        return self._f("obb_dist")(_ptr(a), _ptr(b))

    def obs_distance(self, x10):
        x = np.ascontiguousarray(x10, dtype=np.float64)
        return self._f("obs_distance")(_ptr(x))

    def obs_distance_batch(self, poses_xytht):
        p = np.ascontiguousarray(poses_xytht, dtype=np.float64).reshape(-1, 4)
        out = np.zeros(len(p))
        self._f("obs_distance_batch")(_ptr(p), C.c_int(len(p)), _ptr(out))
        return out

    def simulate(self, state10, ref_x, ref_y, ref_v=None, gb=0, gen_profile=1, vstart=0.0, ref_dir=1, cap=512):
        """Simulation::Simulation(RRT, state, ref, veh, GoalBiased, genProfile, Vstart) on a caller-supplied reference.
        Returns (24-column record, ref_v, stateArray)."""
        st = np.ascontiguousarray(state10, dtype=np.float64)
        x = np.ascontiguousarray(ref_x, dtype=np.float64)
        y = np.ascontiguousarray(ref_y, dtype=np.float64)
        v = np.zeros(len(x)) if ref_v is None else np.array(ref_v, dtype=np.float64)
        out = np.zeros(OUT_STRIDE)
        traj = np.zeros((cap, 10))
        f = self._f("simulate")
        f.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        n = f(_ptr(st), _ptr(x), _ptr(y), _ptr(v), len(x), int(ref_dir), int(gb), int(gen_profile), float(vstart), _ptr(out),
              _ptr(traj), cap)
        return out, v, traj[:min(n, cap)]

    def best_path(self, cap=4096):
        ids = np.zeros(cap, dtype=np.int32)
        n = self._f("best_path")(_ptr(ids), C.c_int(cap))
        return ids[:min(n, cap)]


def check_candidate_lists(planner, samples, heuristic, cand, key, cnt, limit=10):
    """Specification check of top-`limit` feasible candidate lists (rrt/src/rrtplanner.cpp:227-268) that does not
    depend on how equal keys are ordered (std::sort is unstable upstream): for every sample, with all keys and
    feasibility flags recomputed by `planner` (oracle or reference):
      count == min(limit, #feasible); every listed node is feasible and carries exactly its recomputed key;
      listed keys are non-decreasing; no feasible node with a key strictly below the last listed key is missing.
    Returns the number of samples checked."""
    for j in range(len(samples)):
        k_all, feas = planner.keys(samples[j], heuristic[j])
        feas = feas.astype(bool)
        n = int(cnt[j])
        assert n == min(limit, int(feas.sum())), (j, n, int(feas.sum()))
        ids = cand[j, :n]
        assert len(set(ids.tolist())) == n and (cand[j, n:] == -1).all(), (j, cand[j])
        assert feas[ids].all(), (j, ids)
        assert np.array_equal(k_all[ids], key[j, :n]), (j, k_all[ids], key[j, :n])
        assert (np.diff(key[j, :n]) >= 0).all(), (j, key[j, :n])
        if n:
            better = feas & (k_all < key[j, n - 1])
            better[ids] = False
            assert not better.any(), (j, np.where(better)[0])
    return len(samples)


# ---- the synthetic scenes of SURVEY.md §8d -----------------------------------------------------------
def scene_c1_boxes(moving=False):
    """10 static boxes: centre (8+4.7 i, +3 even / -3 odd), theta 0, size_x 4, size_y 8."""
    o = np.zeros((10, 7))
    for i in range(10):
        o[i] = [8 + 4.7 * i, 3.0 if i % 2 == 0 else -3.0, 0.0, 4.0, 8.0, (-1.0 if (moving and i % 2 == 1) else 0.0), 0.0]
    return o


def scene_c3_boxes():
    """1000 oriented boxes, closed-form layout (no RNG)."""
    o = np.zeros((1000, 7))
    for i in range(1000):
        c, r = i % 100, i // 100
        y = 3.0 + 1.5 * (r // 2)
        o[i] = [5.0 + c, y if r % 2 == 0 else -y, (0.1 * i) % np.pi, 2.0, 4.0, 0.0, 0.0]
    return o
