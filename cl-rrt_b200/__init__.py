"""cl-rrt_b200 — Python host binding (ctypes) of the B200-native CL-RRT expansion path.

Thin mirror of include/clrrt.h: the product is libclrrt_b200.so (hand-written sm_100a CUDA behind a C ABI);
this module only marshals numpy arrays.  There is no CPU fallback: if the shared library is missing or no CUDA
device is present, construction raises.

Import name: the directory is called `cl-rrt_b200`, which is not a Python identifier; `import clrrt_b200`
(alias module at the repository root) resolves to this package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLRRT_LIB") or os.path.join(_HERE, "libclrrt_b200.so")  # CLRRT_LIB: kernel-tuning builds only
SORT_LIMIT = 10
RECORD_BYTES = 160
COMM_ID_BYTES = 128


class Vehicle(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("dmax", "ddmax", "Td", "Ta", "amin", "amax", "L", "w", "Lrear", "Lfront", "b", "Vch", "rho", "Kus")]


class Params(C.Structure):
    _fields_ = [("veh", Vehicle), ("sim_dt", C.c_double), ("ctrl_tla", C.c_double), ("ctrl_mindla", C.c_double),
                ("ctrl_dlavmin", C.c_double), ("ctrl_Kp", C.c_double), ("ctrl_Ki", C.c_double),
                ("ref_int", C.c_double), ("ref_mindist", C.c_double), ("ref_res", C.c_double), ("vmax", C.c_double),
                ("ay_road_max", C.c_double), ("Wcost", C.c_double * 5), ("goal", C.c_double * 4),
                ("obs_use_pred", C.c_int32), ("fp32", C.c_int32), ("Cxy", C.c_double * 3), ("lane_shift", C.c_double),
                ("bend", C.c_int32), ("reserved", C.c_int32)]


class RoundStats(C.Structure):
    _fields_ = [("samples", C.c_int32), ("rollouts", C.c_int32), ("nodes_added", C.c_int32),
                ("goal_nodes_added", C.c_int32), ("tree_size", C.c_int32), ("reserved", C.c_int32),
                ("sim_steps", C.c_int64), ("ms_nearest", C.c_float), ("ms_rollout", C.c_float),
                ("ms_prepare", C.c_float), ("ms_append", C.c_float), ("ms_exchange", C.c_float), ("nodes_local", C.c_int32)]


class SeqStats(C.Structure):
    _fields_ = [("iterations", C.c_int64), ("sim_steps", C.c_int64), ("rollouts", C.c_int64), ("windows", C.c_int32),
                ("speculated", C.c_int32), ("nodes_added", C.c_int32), ("tree_size", C.c_int32),
                ("exact_fallbacks", C.c_int32), ("ms_total", C.c_float), ("ms_search", C.c_float), ("ms_prepare", C.c_float),
                ("ms_rollout", C.c_float), ("ms_commit", C.c_float), ("tie_checks_same", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [("fail_collision", C.c_int64), ("fail_acclimit", C.c_int64), ("fail_iterlimit", C.c_int64),
                ("sim_count", C.c_int64), ("rollouts", C.c_int64)]


# numpy views of the C structs clrrt_node / clrrt_rollout / clrrt_obstacle
NODE_DTYPE = np.dtype([("state", "f8", 10), ("ref_front", "f8", 2), ("ref_back", "f8", 2), ("ref_vback", "f8"),
                       ("costE", "f4"), ("costS", "f4"), ("parent", "i4"), ("goal_reached", "i4"), ("n_ref", "i4"),
                       ("kind", "i4"), ("sample", "f8", 2)], align=True)
ROLLOUT_DTYPE = np.dtype([("state", "f8", 10), ("costE", "f8"), ("costS", "f8"), ("ref_back", "f8", 2),
                          ("ref_vback", "f8"), ("trace", "f8"), ("end_reached", "i4"), ("goal_reached", "i4"),
                          ("n_steps", "i4"), ("fail", "i4"), ("n_ref", "i4"), ("idwp0", "i4"), ("tainted", "i4"),
                          ("reserved", "i4")], align=True)
OBSTACLE_DTYPE = np.dtype([(n, "f8") for n in ("cx", "cy", "theta", "size_x", "size_y", "vx", "vy")])


class ClrrtError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load libclrrt_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ClrrtError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, ip, dp = C.c_void_p, C.c_int, C.c_double
    lib.clrrt_default_params.argtypes = [C.POINTER(Params)]
    lib.clrrt_create.argtypes = [C.POINTER(Params), ip, ip, ip, vp, C.POINTER(vp)]
    lib.clrrt_destroy.argtypes = [vp]
    lib.clrrt_last_error.argtypes = [vp]
    lib.clrrt_last_error.restype = C.c_char_p
    lib.clrrt_set_params.argtypes = [vp, C.POINTER(Params)]
    lib.clrrt_set_obstacles.argtypes = [vp, vp, ip]
    lib.clrrt_tree_reset.argtypes = [vp, vp, ip]
    lib.clrrt_tree_size.argtypes = [vp]
    lib.clrrt_tree_truncate.argtypes = [vp, ip]
    lib.clrrt_tree_download.argtypes = [vp, vp, ip, C.POINTER(ip)]
    lib.clrrt_nearest_batch.argtypes = [vp, vp, vp, ip, vp, vp, vp]
    lib.clrrt_propagate_batch.argtypes = [vp, vp, vp, vp, ip, vp, vp, ip]
    lib.clrrt_propagate_batch_ex.argtypes = [vp, vp, vp, vp, ip, vp, vp, ip, vp, ip]
    lib.clrrt_expand_round.argtypes = [vp, vp, vp, ip, C.POINTER(RoundStats)]
    lib.clrrt_expand_round_dev.argtypes = [vp, vp, vp, ip, C.POINTER(RoundStats)]
    lib.clrrt_best_path.argtypes = [vp, vp, ip, C.POINTER(ip)]
    lib.clrrt_counters_get.argtypes = [vp, C.POINTER(Counters)]
    lib.clrrt_set_defer_append.argtypes = [vp, ip]
    lib.clrrt_round_records.argtypes = [vp, C.POINTER(vp), C.POINTER(ip)]
    lib.clrrt_append_records.argtypes = [vp, vp, vp, ip, ip]
    lib.clrrt_set_tuning.argtypes = [vp, ip, ip]
    lib.clrrt_set_tie_mode.argtypes = [vp, ip]
    lib.clrrt_tie_sorts.argtypes = [vp]
    lib.clrrt_tie_sorts.restype = C.c_longlong
    lib.clrrt_set_grid_cell.argtypes = [vp, dp]
    lib.clrrt_set_nearest_mode.argtypes = [vp, ip]
    lib.clrrt_tree_download_range.argtypes = [vp, ip, ip, vp]
    lib.clrrt_tree_download_range_async.argtypes = [vp, ip, ip, vp]
    lib.clrrt_download_wait.argtypes = [vp]
    lib.clrrt_draw_samples.argtypes = [vp, ip, vp, vp]
    lib.clrrt_expand_sequential.argtypes = [vp, vp, vp, ip, ip, C.POINTER(SeqStats)]
    lib.clrrt_comm_unique_id.argtypes = [vp, ip]
    lib.clrrt_comm_init.argtypes = [vp, vp, ip, ip, ip]
    lib.clrrt_comm_attach.argtypes = [vp, vp, ip, ip]
    lib.clrrt_counters_get_global.argtypes = [vp, C.POINTER(Counters)]
    lib.clrrt_tree_digest.argtypes = [vp, ip, ip, vp]
    lib.clrrt_simulate.argtypes = [vp, vp, vp, vp, vp, ip, ip, ip, ip, dp, vp, vp, ip]
    lib.clrrt_simulate_batch.argtypes = [vp, ip, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ip]
    lib.clrrt_collide_batch.argtypes = [vp, vp, ip, vp, vp]
    lib.srand = C.CDLL(None).srand
    _lib = lib
    return lib


def comm_unique_id():
    """ncclGetUniqueId through the library: rank 0 calls it and hands the bytes to the other ranks."""
    buf = (C.c_char * COMM_ID_BYTES)()
    rc = load_library().clrrt_comm_unique_id(buf, COMM_ID_BYTES)
    if rc != 0:
        raise ClrrtError(f"clrrt_comm_unique_id failed ({rc}): NCCL not available")
    return bytes(buf)


def default_params():
    p = Params()
    load_library().clrrt_default_params(C.byref(p))
    return p


def draw_samples(goal, K, seed=None):
    """sampleAroundVehicle + heuristic draw on the C library's rand() (rrt/src/rrtplanner.cpp:133-143, :187-201)."""
    lib = load_library()
    if seed is not None:
        C.CDLL(None).srand(C.c_uint(seed))
    g = (C.c_double * 4)(*[float(x) for x in goal])
    s = np.zeros((K, 2))
    h = np.zeros(K, np.uint8)
    rc = lib.clrrt_draw_samples(g, K, s.ctypes.data, h.ctypes.data)
    if rc != 0:
        raise ClrrtError(f"clrrt_draw_samples failed ({rc})")
    return s, h


def root_node(car_state6):
    """MyRRT::addInitialNode (rrt/src/rrtplanner.cpp:21-37) for a car state [x,y,theta,delta,v,a] already in the
    car frame (x=y=theta=0, rrt/src/transformations.cpp:143-147): 10-point reference (0,0)->(1,0) built by
    LinearSpacedVector's accumulation, ref.v = state[4]."""
    n = np.zeros(1, NODE_DTYPE)
    s = np.zeros(10)
    s[3:6] = np.asarray(car_state6, float)[3:6]
    n["state"][0] = s
    N = int(np.floor(np.sqrt(1.0 ** 2 + 0.0 ** 2) / 0.1))
    h = (1.0 - 0.0) / float(N - 1)
    val = 0.0
    for _ in range(N - 1):
        val += h
    n["ref_front"][0] = (0.0, 0.0)
    n["ref_back"][0] = (val, 0.0)
    n["ref_vback"][0] = s[4]
    n["parent"][0] = -1
    n["n_ref"][0] = N
    return n


class Planner:
    """One context = one GPU (clrrt_create)."""

    def __init__(self, params=None, device=0, tree_capacity=1 << 16, max_round=1 << 12, stream=None):
        self.lib = load_library()
        self.params = params if params is not None else default_params()
        h = C.c_void_p()
        rc = self.lib.clrrt_create(C.byref(self.params), device, tree_capacity, max_round, stream, C.byref(h))
        self.h = h
        if rc != 0:
            msg = self.lib.clrrt_last_error(h).decode() if h else "no CUDA device / bad arguments"
            if h:
                self.lib.clrrt_destroy(h)
                self.h = None
            raise ClrrtError(f"clrrt_create failed ({rc}): {msg}")
        self.max_round = max_round
        self.tree_capacity = tree_capacity

    def close(self):
        if getattr(self, "h", None):
            self.lib.clrrt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise ClrrtError(f"clrrt error {rc}: {self.lib.clrrt_last_error(self.h).decode()}")

    # ---- configuration --------------------------------------------------------------------------------
    def set_params(self, params=None):
        if params is not None:
            self.params = params
        self._ck(self.lib.clrrt_set_params(self.h, C.byref(self.params)))

    def set_query(self, car_state6, goal, vmax):
        """The per-query globals of MotionPlanner::planMotion (rrt/src/motionplanner.cpp:16-17)."""
        p = self.params
        v = abs(float(car_state6[4]))
        a, b = v * p.ref_int, p.ref_mindist
        p.ref_res = b if a < b else a  # std::max(abs(v)*ref_int, ref_mindist), rrt/src/controller.cpp:20
        p.vmax = float(vmax)
        for i in range(4):
            p.goal[i] = float(goal[i])
        self.set_params()

    def set_road(self, bend, Cxy3=(0.0, 0.0, 0.0), lane_shift=0.0):
        """MotionRequest.bend / Cxy / laneShifts[0] (rrt/src/motionplanner.cpp:23): lane-deviation cost per sim step."""
        p = self.params
        p.bend = 1 if bend else 0
        for i in range(3):
            p.Cxy[i] = float(Cxy3[i])
        p.lane_shift = float(lane_shift)
        self.set_params()

    def set_obstacles(self, obstacles):
        o = np.ascontiguousarray(obstacles, dtype=np.float64).reshape(-1, 7)
        self._ck(self.lib.clrrt_set_obstacles(self.h, o.ctypes.data if len(o) else None, len(o)))

    def set_tie_mode(self, mode=1):
        """1 (default): K = 1 searches order equal keys as the reference's std::sort does; 0: lower node id first."""
        self._ck(self.lib.clrrt_set_tie_mode(self.h, mode))

    def tie_sorts(self):
        return int(self.lib.clrrt_tie_sorts(self.h))

    def set_tuning(self, refill_min=8, blocks_per_sm=0):
        self._ck(self.lib.clrrt_set_tuning(self.h, refill_min, blocks_per_sm))

    def set_nearest_mode(self, mode):
        self._ck(self.lib.clrrt_set_nearest_mode(self.h, int(mode)))

    def set_grid_cell(self, metres):
        self._ck(self.lib.clrrt_set_grid_cell(self.h, float(metres)))

    # ---- tree -----------------------------------------------------------------------------------------
    def tree_reset(self, nodes):
        nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
        self._ck(self.lib.clrrt_tree_reset(self.h, nodes.ctypes.data, len(nodes)))

    def tree_reset_records(self, rec20):
        """From the 20-double record format of the CPU libraries (tests/cpulib.py)."""
        r = np.asarray(rec20, float).reshape(-1, 20)
        n = np.zeros(len(r), NODE_DTYPE)
        n["state"] = r[:, :10]
        n["ref_front"] = r[:, 10:12]
        n["ref_back"] = r[:, 12:14]
        n["ref_vback"] = r[:, 14]
        n["costE"] = r[:, 15].astype(np.float32)
        n["costS"] = r[:, 16].astype(np.float32)
        n["parent"] = r[:, 17].astype(np.int32)
        n["goal_reached"] = r[:, 18].astype(np.int32)
        n["n_ref"] = r[:, 19].astype(np.int32)
        self.tree_reset(n)

    def tree_size(self):
        return self.lib.clrrt_tree_size(self.h)

    def tree_truncate(self, n):
        self._ck(self.lib.clrrt_tree_truncate(self.h, n))

    def tree_download(self):
        n = self.tree_size()
        out = np.zeros(n, NODE_DTYPE)
        got = C.c_int(0)
        self._ck(self.lib.clrrt_tree_download(self.h, out.ctypes.data, n, C.byref(got)))
        return out[:got.value]

    def tree_download_range(self, first, count, out=None):
        """`out`: optional caller buffer (e.g. a view of pinned memory) of at least `count` NODE_DTYPE records."""
        out = np.zeros(count, NODE_DTYPE) if out is None else out[:count]
        self._ck(self.lib.clrrt_tree_download_range(self.h, first, count, out.ctypes.data))
        return out

    def tree_download_range_async(self, first, count, out):
        """Stage nodes [first, first + count) now, copy them to `out` (a view of PINNED memory) beside the next round;
        download_wait() before reading `out`."""
        self._ck(self.lib.clrrt_tree_download_range_async(self.h, first, count, out.ctypes.data))
        return out[:count]

    def download_wait(self):
        self._ck(self.lib.clrrt_download_wait(self.h))

    def tree_download_records(self):
        t = self.tree_download()
        r = np.zeros((len(t), 20))
        r[:, :10] = t["state"]
        r[:, 10:12] = t["ref_front"]
        r[:, 12:14] = t["ref_back"]
        r[:, 14] = t["ref_vback"]
        r[:, 15] = t["costE"]
        r[:, 16] = t["costS"]
        r[:, 17] = t["parent"]
        r[:, 18] = t["goal_reached"]
        r[:, 19] = t["n_ref"]
        return r

    # ---- batched primitives ---------------------------------------------------------------------------
    def nearest_batch(self, samples, heuristic):
        s = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, 2)
        h = np.ascontiguousarray(heuristic, dtype=np.uint8)
        K = len(s)
        cand = np.zeros((K, SORT_LIMIT), np.int32)
        key = np.zeros((K, SORT_LIMIT), np.float32)
        cnt = np.zeros(K, np.int32)
        self._ck(self.lib.clrrt_nearest_batch(self.h, s.ctypes.data, h.ctypes.data, K, cand.ctypes.data,
                                              key.ctypes.data, cnt.ctypes.data))
        return cand, key, cnt

    def propagate_batch(self, parent, samples, goal_biased=None, traj_stride=0, ref_stride=0):
        par = np.ascontiguousarray(parent, dtype=np.int32)
        s = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, 2)
        M = len(par)
        gb = None if goal_biased is None else np.ascontiguousarray(goal_biased, dtype=np.uint8)
        out = np.zeros(M, ROLLOUT_DTYPE)
        traj = np.zeros((M, traj_stride, 10)) if traj_stride else None
        ref = np.zeros((M, ref_stride, 3)) if ref_stride else None
        self._ck(self.lib.clrrt_propagate_batch_ex(self.h, par.ctypes.data, s.ctypes.data,
                                                   None if gb is None else gb.ctypes.data, M, out.ctypes.data,
                                                   None if traj is None else traj.ctypes.data, traj_stride,
                                                   None if ref is None else ref.ctypes.data, ref_stride))
        if ref_stride:
            return out, traj, ref
        return (out, traj) if traj_stride else out

    def simulate(self, state10, ref_x, ref_y, ref_v=None, goal_biased=False, gen_profile=True, vstart=0.0, ref_dir=1,
                 traj_stride=0):
        """Simulation::Simulation(RRT, state, ref, veh, GoalBiased, genProfile, Vstart) (rrt/include/rrt/simulation.h:18-19)
        on a caller-owned reference.  Returns (rollout record, ref_v [filled when gen_profile], traj or None)."""
        st = np.ascontiguousarray(state10, dtype=np.float64)
        x = np.ascontiguousarray(ref_x, dtype=np.float64)
        y = np.ascontiguousarray(ref_y, dtype=np.float64)
        v = np.zeros(len(x)) if ref_v is None else np.array(ref_v, dtype=np.float64)
        out = np.zeros(1, ROLLOUT_DTYPE)
        traj = np.zeros((traj_stride, 10)) if traj_stride else None
        self._ck(self.lib.clrrt_simulate(self.h, st.ctypes.data, x.ctypes.data, y.ctypes.data, v.ctypes.data, len(x), int(ref_dir),
                                         int(bool(goal_biased)), int(bool(gen_profile)), float(vstart), out.ctypes.data,
                                         None if traj is None else traj.ctypes.data, traj_stride))
        return out[0], v, traj

    def simulate_batch(self, states, refs, goal_biased=None, gen_profile=None, vstart=None, ref_dir=None):
        """M rollouts with ragged references: refs = list of (x, y[, v]) arrays.  Returns (records, list of ref_v)."""
        M = len(refs)
        st = np.ascontiguousarray(states, dtype=np.float64).reshape(M, 10)
        off = np.zeros(M + 1, np.int32)
        off[1:] = np.cumsum([len(r[0]) for r in refs])
        x = np.concatenate([np.asarray(r[0], float) for r in refs])
        y = np.concatenate([np.asarray(r[1], float) for r in refs])
        v = np.concatenate([np.asarray(r[2], float) if len(r) > 2 and r[2] is not None else np.zeros(len(r[0])) for r in refs])
        gb = np.zeros(M, np.uint8) if goal_biased is None else np.ascontiguousarray(goal_biased, dtype=np.uint8)
        gp = np.ones(M, np.uint8) if gen_profile is None else np.ascontiguousarray(gen_profile, dtype=np.uint8)
        vs = np.zeros(M) if vstart is None else np.ascontiguousarray(vstart, dtype=np.float64)
        dr = np.ones(M, np.int32) if ref_dir is None else np.ascontiguousarray(ref_dir, dtype=np.int32)
        out = np.zeros(M, ROLLOUT_DTYPE)
        self._ck(self.lib.clrrt_simulate_batch(self.h, M, st.ctypes.data, off.ctypes.data, x.ctypes.data, y.ctypes.data,
                                               v.ctypes.data, gb.ctypes.data, gp.ctypes.data, vs.ctypes.data, dr.ctypes.data,
                                               out.ctypes.data, None, 0))
        return out, [v[off[i]:off[i + 1]] for i in range(M)]

    def collide_batch(self, poses_xytht, want_distance=False):
        """checkObsDistance for n poses (x, y, theta, t): verdict (1 = collision) from the verdict-only path and, on
        request, the reference's pseudo-distance from the exact path."""
        p = np.ascontiguousarray(poses_xytht, dtype=np.float64).reshape(-1, 4)
        v = np.zeros(len(p), np.int32)
        d = np.zeros(len(p)) if want_distance else None
        self._ck(self.lib.clrrt_collide_batch(self.h, p.ctypes.data, len(p), v.ctypes.data, None if d is None else d.ctypes.data))
        return (v, d) if want_distance else v

    def expand_round(self, samples, heuristic):
        s = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, 2)
        h = np.ascontiguousarray(heuristic, dtype=np.uint8)
        st = RoundStats()
        self._ck(self.lib.clrrt_expand_round(self.h, s.ctypes.data, h.ctypes.data, len(s), C.byref(st)))
        return st

    def expand_sequential(self, samples, heuristic, window=0):
        """len(samples) consecutive expandTree calls (the reference's sequential algorithm), executed as speculative
        windows on the device (clrrt_expand_sequential).  window: samples in flight (0 = adaptive)."""
        s = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, 2)
        h = np.ascontiguousarray(heuristic, dtype=np.uint8)
        st = SeqStats()
        self._ck(self.lib.clrrt_expand_sequential(self.h, s.ctypes.data, h.ctypes.data, len(s), int(window), C.byref(st)))
        return st

    def expand_round_dev(self, d_samples_ptr, d_heur_ptr, K):
        st = RoundStats()
        self._ck(self.lib.clrrt_expand_round_dev(self.h, d_samples_ptr, d_heur_ptr, K, C.byref(st)))
        return st

    def best_path(self, cap=4096):
        ids = np.zeros(cap, np.int32)
        n = C.c_int(0)
        self._ck(self.lib.clrrt_best_path(self.h, ids.ctypes.data, cap, C.byref(n)))
        return ids[:min(n.value, cap)]

    def counters(self):
        c = Counters()
        self._ck(self.lib.clrrt_counters_get(self.h, C.byref(c)))
        return dict(fail_collision=c.fail_collision, fail_acclimit=c.fail_acclimit,
                    fail_iterlimit=c.fail_iterlimit, sim_count=c.sim_count, rollouts=c.rollouts)

    # ---- multi-GPU inside the library (clrrt_comm_*) ------------------------------------------------------
    def comm_init(self, unique_id, rank, world):
        """ncclCommInitRank on this context's device (collective).  unique_id: the 128 bytes of comm_unique_id()."""
        buf = (C.c_char * COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.clrrt_comm_init(self.h, buf, COMM_ID_BYTES, rank, world))

    def counters_global(self):
        c = Counters()
        self._ck(self.lib.clrrt_counters_get_global(self.h, C.byref(c)))
        return dict(fail_collision=c.fail_collision, fail_acclimit=c.fail_acclimit,
                    fail_iterlimit=c.fail_iterlimit, sim_count=c.sim_count, rollouts=c.rollouts)

    def tree_digest(self, first=0, count=None):
        """128-bit order-sensitive digest of tree nodes [first, first + count) as a hex string."""
        count = self.tree_size() - first if count is None else count
        out = (C.c_uint64 * 2)()
        self._ck(self.lib.clrrt_tree_digest(self.h, first, count, out))
        return f"{out[0]:016x}{out[1]:016x}"

    # ---- multi-GPU exchange, lower-level form -------------------------------------------------------------
    def set_defer_append(self, defer):
        self._ck(self.lib.clrrt_set_defer_append(self.h, int(defer)))

    def round_records(self):
        ptr, n = C.c_void_p(), C.c_int(0)
        self._ck(self.lib.clrrt_round_records(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def append_records(self, d_records_ptr, counts, stride_records):
        c = np.ascontiguousarray(counts, dtype=np.int32)
        self._ck(self.lib.clrrt_append_records(self.h, d_records_ptr, c.ctypes.data, len(c), stride_records))


def rollouts_as_table(out):
    """ROLLOUT_DTYPE array -> the 24-column table used by the CPU libraries (tests/cpulib.py)."""
    t = np.zeros((len(out), 24))
    t[:, :10] = out["state"]
    t[:, 10] = out["costE"]
    t[:, 11] = out["costS"]
    t[:, 12] = out["end_reached"]
    t[:, 13] = out["goal_reached"]
    t[:, 14] = out["n_steps"]
    t[:, 15] = out["fail"]
    t[:, 16] = out["n_ref"]
    t[:, 17:19] = out["ref_back"]
    t[:, 19] = out["ref_vback"]
    t[:, 20] = out["tainted"]
    t[:, 21] = out["trace"]
    t[:, 22] = out["idwp0"]
    return t
