#!/usr/bin/env python
"""bench.py — CL-RRT tree-expansion throughput on B200 (contract in the task statement, tier section ④).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU code (oracle/_ref)

A *step* is one expansion round of config C3: 65 536 samples (per GPU) against the 4096-node oracle-grown
tree snapshot tests/golden/c3_snapshot.npz in the dense urban scene (1000 oriented boxes, car at rest, goal 100 m ahead):
candidate search, closed-loop rollouts in candidate order until the first success, goal-biased rollouts, ordered append
(+ the node all-gather inside the library when N > 1).  The tree is truncated back to the snapshot after every round so
that every step does the same work.  Metric: closed-loop sim steps per second (one sim step = one iteration of
rrt/src/simulation.cpp:58), whole job.  Sub-records of the same JSON line: `c4` (2^20 samples per round in total, strong
scaling, with the device digest of the grown tree), `fp32`, `c2`, `c5`, `query_200ms` (snapshot rounds) and
`query_200ms_k1` (the reference's sequential algorithm), `sustained`.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "closed_loop_sim_steps_per_s"
UNIT = "sim-steps/s"
K_ROUND = 65536
TREE_SNAPSHOT = 4096
CAR = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0)   # SURVEY.md §8d "Common": car at rest unless stated
GOAL = (100.0, 0.0, 0.0, 0.0)
VMAX = 5.0
SNAPSHOT_FILE = os.path.join(ROOT, "tests", "golden", "c3_snapshot.npz")

C1_CAR = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
C1_GOAL = (50.0, 0.0, 0.0, 0.0)
C2_CAR = (0.0, 0.0, 0.0, 0.0, 3.0, 0.0)
C4_TOTAL = 1 << 20


def workload_config(world, fp32=False):
    """The `config` object of the JSON line: identical for both arms."""
    return {"workload": "C3 dense urban scene: 1000 OBB obstacles, 65536 samples/round/GPU against the 4096-node oracle-grown tree "
                        "snapshot (tests/golden/c3_snapshot.npz), car at rest, goal 100 m ahead, "
                        + ("fp32 mode" if fp32 else "fp64 parity mode"),
            "samples_per_round_per_gpu": K_ROUND, "tree_nodes": TREE_SNAPSHOT, "obstacles": 1000,
            "parallelism": (f"samples sharded over {world} GPU(s), tree replicated, per-round node all-gather (NCCL) inside the library"
                            if world > 1 else "1 GPU"),
            "l2_policy": "per-round inputs+outputs (samples, candidate lists, prepared rollouts, staging SoA, records: >300 MB) are "
                         "rewritten every round, larger than the 126 MB L2; the kernel is compute-bound"}


def scene_c3_boxes():
    """closed-form layout, no RNG."""
    o = np.zeros((1000, 7))
    for i in range(1000):
        c, r = i % 100, i // 100
        y = 3.0 + 1.5 * (r // 2)
        o[i] = [5.0 + c, y if r % 2 == 0 else -y, (0.1 * i) % np.pi, 2.0, 4.0, 0.0, 0.0]
    return o


def scene_c1_boxes():
    """10 static boxes, centre (8 + 4.7 i, +3 even / -3 odd), size_x 4, size_y 8."""
    o = np.zeros((10, 7))
    for i in range(10):
        o[i] = [8 + 4.7 * i, 3.0 if i % 2 == 0 else -3.0, 0.0, 4.0, 8.0, 0.0, 0.0]
    return o


def load_snapshot():
    g = np.load(SNAPSHOT_FILE)
    assert tuple(g["car"]) == CAR and tuple(g["goal"]) == GOAL
    return np.ascontiguousarray(g["tree"][:TREE_SNAPSHOT])


def draw_global_samples(goal, n, seed):
    """sampleAroundVehicle + the heuristic draw on the C library's rand() (rrt/src/rrtplanner.cpp:133-143, :187-201): the same
    expressions in clrrt_draw_samples and in the oracle; this is the product's (no oracle on the product path)."""
    import clrrt_b200 as clrrt
    return clrrt.draw_samples(goal, n, seed=seed)


def build_workload(pl, clrrt, rank, world):
    """C3 on a planner: scene, the golden tree snapshot, this rank's shard of the global sample stream (used by the
    diagnostic scripts under scripts/ as well)."""
    boxes = scene_c3_boxes()
    pl.set_query(CAR, GOAL, VMAX)
    pl.set_obstacles(boxes)
    pl.tree_reset_records(load_snapshot())
    gs, gh = draw_global_samples(GOAL, K_ROUND * world, seed=2)
    return boxes, gs[rank * K_ROUND:(rank + 1) * K_ROUND].copy(), gh[rank * K_ROUND:(rank + 1) * K_ROUND].copy()


def _as_cuda_tensor(ptr, nbytes, device):
    """Zero-copy torch view of a device buffer owned by the library (scripts/multirank_check.py: the lower-level exchange)."""
    import torch

    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    return torch.as_tensor(h, device=torch.device("cuda", device))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe): one nvidia-smi process
    sampling every 100 ms from before the warm-up until after the last timed step; rows carry their own timestamps and
    only those inside [mark_begin, mark_end] are summarised."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.t0, self.t1 = None, None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def summary(self):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ""
        rows = []
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[3:7]))
            except Exception:
                continue
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi gave no samples"]}
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.05 <= r[0] <= (self.t1 or 1e18) + 0.05] or rows
        sm = sorted(r[1] for r in inside)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3][k].lower().startswith("active") for r in inside)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": inside[0][2], "reasons": reasons, "samples": len(inside),
                "samples_total": len(rows)}


# ---------------------------------------------------------------------------------------------------------------------
# sub-records (this repo's arm)
# ---------------------------------------------------------------------------------------------------------------------
def query_200ms_rounds(clrrt, device, K=16384, budget_ms=200.0):
    """Tree nodes grown in 200 ms by SNAPSHOT ROUNDS of K samples (config C1 scene).  This is the batched formulation — a
    different algorithm from the reference's sequential loop for K > 1; the like-for-like figure is
    query_200ms_k1.  Samples are drawn on the host with the reference's expressions; the wall clock covers drawing, the
    host->device copy and the round."""
    pl = clrrt.Planner(device=device, tree_capacity=(1 << 20) + 2 * K, max_round=K)
    pl.set_query(C1_CAR, C1_GOAL, VMAX)
    pl.set_obstacles(scene_c1_boxes())
    pl.tree_reset(clrrt.root_node(C1_CAR))
    s, h = clrrt.draw_samples(C1_GOAL, K, seed=1)
    pl.expand_round(s, h)  # warm-up (first launch, lazy module load), then start over
    pl.tree_reset(clrrt.root_node(C1_CAR))
    clrrt.draw_samples(C1_GOAL, 1, seed=1)
    rounds = steps = rollouts = 0
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=1)
    t0 = time.perf_counter()
    nxt = pool.submit(clrrt.draw_samples, C1_GOAL, K)
    full = False
    while (time.perf_counter() - t0) * 1e3 < budget_ms and not full:
        s, h = nxt.result()
        nxt = pool.submit(clrrt.draw_samples, C1_GOAL, K)
        try:
            st = pl.expand_round(s, h)
        except clrrt.ClrrtError:
            full = True
            break
        rounds += 1
        steps += st.sim_steps
        rollouts += st.rollouts
    wall = (time.perf_counter() - t0) * 1e3
    nxt.result()
    pool.shutdown()
    nodes = pl.tree_size()
    path = len(pl.best_path())
    pl.close()
    return {"nodes": int(nodes), "rounds": rounds, "samples_per_round": K, "sim_steps": int(steps), "rollouts": int(rollouts),
            "wall_ms": wall, "budget_ms": budget_ms, "best_path_nodes": int(path), "tree_full": full,
            "algorithm": f"snapshot rounds of {K} samples (NOT the reference's sequential loop: compare query_200ms_k1)",
            "scene": "C1: straight road, goal 50 m ahead, 10 static boxes, Prius parameters"}


def query_200ms_k1(clrrt, device, budget_ms=200.0, chunk=32):
    """The reference's own algorithm — one sample per expandTree, every sample seeing the nodes of the previous ones
    (rrt/src/motionplanner.cpp:39-43) — for 200 ms of wall clock on the C1 scene: clrrt_expand_sequential, `chunk`
    iterations per call with several samples in flight on the device, timer polled between calls.  Same draws (srand(1))
    as the reference arm's loop, so the first N nodes are the same nodes."""
    pl = clrrt.Planner(device=device, tree_capacity=1 << 16, max_round=256)
    pl.set_query(C1_CAR, C1_GOAL, VMAX)
    pl.set_obstacles(scene_c1_boxes())
    pl.tree_reset(clrrt.root_node(C1_CAR))
    s, h = clrrt.draw_samples(C1_GOAL, 64, seed=1)
    pl.expand_sequential(s, h)  # warm-up, then start over with the same stream
    pl.tree_reset(clrrt.root_node(C1_CAR))
    clrrt.draw_samples(C1_GOAL, 1, seed=1)
    ctypes.CDLL(None).srand(ctypes.c_uint(1))
    it = steps = windows = spec = 0
    dev = [0.0, 0.0, 0.0, 0.0]
    t0 = time.perf_counter()
    while (time.perf_counter() - t0) * 1e3 < budget_ms:
        s, h = clrrt.draw_samples(C1_GOAL, chunk)
        st = pl.expand_sequential(s, h)
        it += chunk
        steps += st.sim_steps
        windows += st.windows
        spec += st.speculated
        for k, v in enumerate((st.ms_search, st.ms_prepare, st.ms_rollout, st.ms_commit)):
            dev[k] += v
    wall = (time.perf_counter() - t0) * 1e3
    nodes = pl.tree_size()
    path = len(pl.best_path())
    pl.close()
    return {"nodes": int(nodes), "iterations": it, "sim_steps": int(steps), "wall_ms": wall, "budget_ms": budget_ms,
            "windows": windows, "samples_speculated": spec, "best_path_nodes": int(path),
            "device_ms": {"search": dev[0], "prepare": dev[1], "rollout_kernel": dev[2], "select_tie_commit": dev[3]},
            "algorithm": "the reference's sequential expandTree loop (tree identical to the reference's for the same number of "
                         "iterations), speculative windows on the device",
            "scene": "C1: straight road, goal 50 m ahead, 10 static boxes, Prius parameters"}


def c2_batch(clrrt, device, reps=5):
    """Config C2: 4096 sampled rollouts (top-1 feasible parent, explore key) on the 10-box scene from a tree the product grows
    with 60 sequential iterations (srand(1), v0 = 3) — sim steps per second through clrrt_propagate_batch with host buffers."""
    pl = clrrt.Planner(device=device, tree_capacity=1 << 14, max_round=4096)
    pl.set_query(C2_CAR, C1_GOAL, VMAX)
    pl.set_obstacles(scene_c1_boxes())
    pl.tree_reset(clrrt.root_node(C2_CAR))
    s, h = clrrt.draw_samples(C1_GOAL, 60, seed=1)
    pl.expand_sequential(s, h)
    smp, heu = clrrt.draw_samples(C1_GOAL, 4096)
    cand, _, cnt = pl.nearest_batch(smp, np.zeros(4096, np.uint8))
    ok = cnt > 0
    par, sm = cand[ok, 0], smp[ok]
    pl.propagate_batch(par, sm)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = pl.propagate_batch(par, sm)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    steps = int(out["n_steps"].sum())
    t0 = time.perf_counter()
    st = pl.expand_round(smp, heu)
    dt_round = time.perf_counter() - t0
    n_tree = pl.tree_size() - st.nodes_added
    pl.close()
    return {"rollouts": int(ok.sum()), "sim_steps": steps, "ms": best * 1e3, "value": steps / best, "unit": UNIT,
            "tree_nodes": int(n_tree), "round_4096": {"ms": dt_round * 1e3, "sim_steps": int(st.sim_steps), "value": st.sim_steps / dt_round},
            "what": "C2: 4096 top-1 rollouts, 10 boxes, fp64 parity mode, host buffers in and out (clrrt_propagate_batch); round_4096 = "
                    "one clrrt_expand_round of the same 4096 samples (<= 10 candidates each)"}


def c5_loop(clrrt, device):
    """Config C5: the recorded 100-query receding-horizon loop (tests/golden/g5_replan.npz: world states, goals and moving
    obstacles of the reference's own run, commit_path = true, 100 expandTree iterations per query) through the C++ host
    facade, sequential mode: wall time per query and nodes per query.  tests/test_gpu_replan.py checks the same loop against
    the reference query by query."""
    path = os.path.join(ROOT, "tests", "golden", "g5_replan.npz")
    if not os.path.exists(path):
        return None
    from test_gpu_replan import HostPlanner
    g = np.load(path)
    iters, nq = int(g["iters"]), int(g["world"].shape[0])
    hp = HostPlanner(samples_per_round=1, commit_path=True, device=device)
    ctypes.CDLL(None).srand(ctypes.c_uint(1))
    hp.query(g["world"][0], g["goal"][0], g["obstacles"][0], 4)  # warm-up (module load), then a fresh planner
    hp.close()
    hp = HostPlanner(samples_per_round=1, commit_path=True, device=device)
    ctypes.CDLL(None).srand(ctypes.c_uint(1))
    nodes = steps = 0
    same = 0
    t0 = time.perf_counter()
    for q in range(nq):
        sizes, cost, cnt = hp.query(g["world"][q], g["goal"][q], g["obstacles"][q], iters)
        nodes += int(sizes[1])
        steps += int(cnt.sim_count)
        same += int(sizes[:3].tolist() == g["sizes"][q].tolist() and cnt.sim_count == int(g["sim_steps"][q]))
    wall = time.perf_counter() - t0
    hp.close()
    return {"queries": nq, "iterations_per_query": iters, "wall_ms_per_query": 1e3 * wall / nq, "nodes_per_query": nodes / nq,
            "sim_steps_per_query": steps / nq, "value": steps / wall, "unit": UNIT,
            "queries_equal_to_reference_golden": same,
            "what": "C5: 100 consecutive planMotion queries, tree re-initialised from the previous best path, moving obstacles; "
                    "C++ host facade (libclrrt_host.so), the reference's sequential algorithm"}


def c4_strong(clrrt, torch, dist, rank, world, local, stream, steps=5, warmup=2):
    """Config C4: 2^20 samples per round IN TOTAL, sharded over the ranks (strong scaling), the 10 boxes of C1/C2, 4096-node
    tree snapshot grown by the product itself (deterministic: srand(1) rounds of 8192), node all-gather inside the library.
    tree_digest: device digest of the tree after one round — identical for every world size."""
    K = C4_TOTAL // world
    pl = clrrt.Planner(device=local, tree_capacity=TREE_SNAPSHOT + 2 * C4_TOTAL + 1024, max_round=K, stream=stream.cuda_stream)
    if world > 1:
        uid = [clrrt.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        pl.comm_init(uid[0], rank, world)
    pl.set_query(C2_CAR, C1_GOAL, VMAX)
    pl.set_obstacles(scene_c1_boxes())
    pl.tree_reset(clrrt.root_node(C2_CAR))
    s, h = clrrt.draw_samples(C1_GOAL, 8192 * 12, seed=1)
    i = 0
    while pl.tree_size() < TREE_SNAPSHOT and i < 12:   # every rank grows the same snapshot from the same global samples
        gs, gh = s[i * 8192:(i + 1) * 8192], h[i * 8192:(i + 1) * 8192]
        lo, hi = rank * 8192 // world, (rank + 1) * 8192 // world
        pl.expand_round(gs[lo:hi] if world > 1 else gs, gh[lo:hi] if world > 1 else gh)
        i += 1
    pl.tree_truncate(TREE_SNAPSHOT)
    n0 = pl.tree_size()
    gs, gh = clrrt.draw_samples(C1_GOAL, C4_TOTAL, seed=2)
    d_s = torch.from_numpy(gs[rank * K:(rank + 1) * K].copy()).cuda()
    d_h = torch.from_numpy(gh[rank * K:(rank + 1) * K].copy()).cuda()
    digest = None
    for _ in range(warmup):
        st = pl.expand_round_dev(d_s.data_ptr(), d_h.data_ptr(), K)
        if digest is None:
            digest, grown = pl.tree_digest(), pl.tree_size()
        pl.tree_truncate(n0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    tot = 0
    ms_ex = ms_roll = ms_nn = 0.0
    for _ in range(steps):
        st = pl.expand_round_dev(d_s.data_ptr(), d_h.data_ptr(), K)
        tot += st.sim_steps
        ms_ex += st.ms_exchange
        ms_roll += st.ms_rollout
        ms_nn += st.ms_nearest
        pl.tree_truncate(n0)
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    digests = [digest]
    if world > 1:
        t = torch.tensor([ms, float(tot)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, tot = float(tmax[0]), int(t[1])
        digests = [None] * world
        dist.all_gather_object(digests, digest)
    pl.close()
    return {"workload": f"C4: 2^20 samples/round in total ({K} per GPU), 10 OBB obstacles, 4096-node snapshot, fp64 parity mode",
            "scaling": "strong", "n_gpus": world, "steps": steps, "warmup": warmup, "value": tot / (ms * 1e-3), "unit": UNIT,
            "ms_per_step": ms / steps, "sim_steps_per_round": tot / steps, "nodes_after_one_round": int(grown),
            "tree_digest": digest, "digest_equal_on_all_ranks": all(d == digest for d in digests),
            "ms_per_round_rank0": {"nearest": ms_nn / steps, "rollout": ms_roll / steps, "exchange_and_append": ms_ex / steps}}


def cpu_baseline_sample(boxes, tree_records, samples, heur, budget_s=12.0, kind=None):
    """The reference's own CPU code (oracle/_ref when present, else the C restatement) on a bounded sample of the
    same workload: top-1 candidate rollouts, batches of 32 until the time budget is spent."""
    from cpulib import CpuPlanner, ref_available
    kind = kind or ("ref" if ref_available(False) else "oracle")
    cpu = CpuPlanner(kind)
    cpu.set_obstacles(boxes)
    cpu.tree_init(CAR, GOAL, VMAX)
    cpu.tree_import(tree_records)
    steps = rollouts = 0
    secs = 0.0
    t_nn = 0.0
    j = 0
    while secs + t_nn < budget_s and j + 32 <= len(samples):
        cand, _, cnt = cpu.nearest_batch(samples[j:j + 32], heur[j:j + 32])
        t_nn += cpu.last_seconds
        ok = cnt > 0
        out = cpu.rollout_batch(cand[ok, 0], samples[j:j + 32][ok])
        secs += cpu.last_seconds
        steps += int(out[:, 14].sum())
        rollouts += int(ok.sum())
        j += 32
    return {"value": steps / max(secs + t_nn, 1e-9), "unit": UNIT, "cores": 1,
            "kind": "reference" if kind.startswith("ref") else "port",
            "sample": f"{rollouts} top-1 candidate rollouts ({steps} sim steps) + their candidate search for {j} samples "
                      f"of the same round, {secs + t_nn:.1f} s on 1 host core ({'oracle/_ref (unmodified reference sources)' if kind.startswith('ref') else 'oracle C port'}, -O3 -DNDEBUG, ROS logging compiled out)",
            "rollouts_per_s": rollouts / max(secs + t_nn, 1e-9)}


def work_profile(boxes, tree_records, samples, heur, n=32):
    """R (reference points scanned per waypoint search) and A (SAT axes per box pair) as the reference executes
    this workload, counted by the oracle port on a few rollouts (SURVEY.md §8d uses them in the FLOP count)."""
    from cpulib import CpuPlanner
    orc = CpuPlanner("oracle")
    orc.set_obstacles(boxes)
    orc.tree_init(CAR, GOAL, VMAX)
    orc.tree_import(tree_records)
    c0 = (ctypes.c_long * 3)()
    orc._f("work_counters")(c0)
    cand, _, cnt = orc.nearest_batch(samples[:n], heur[:n])
    ok = cnt > 0
    out = orc.rollout_batch(cand[ok, 0], samples[:n][ok])
    c1 = (ctypes.c_long * 3)()
    orc._f("work_counters")(c1)
    steps = max(1, int(out[:, 14].sum()))
    sat_calls, sat_axes, scanned = (c1[i] - c0[i] for i in range(3))
    return scanned / steps, (sat_axes / sat_calls if sat_calls else 1.0)


def algorithmic_flop_per_step(n_obs, R=1.0, A=1.17):
    """F_core 166 + waypoint scan 6R + collision 38 + n_obs (42 + 40 A)."""
    return 166.0 + 6.0 * R + 38.0 + n_obs * (42.0 + 40.0 * A)


def executed_roofline(kernel_ms, sm_mhz, n_sms, fp32, peaks, per_launch_steps, hbm_bytes):
    """Roofline of the round kernel from the work it EXECUTES: the FP64 (fp32 mode: FP32) thread instructions the kernel
    retires per launch — counted once by ncu on this same workload and committed in profiles/r02_rollout_ops.json
    (scripts/rollout_ops.py; the counts are a property of the workload, not of the clock) — times 1 FLOP (add, mul) or
    2 (fma), divided by the kernel's duration measured LIVE in this run, against the pipe's peak at the SM clock sampled in
    this run: FP64 = SMs x 64 lanes x 2 x f, FP32 = SMs x 128 lanes x 2 x f."""
    try:
        ops = json.load(open(os.path.join(ROOT, "profiles", "r02_rollout_ops.json")))["fp32" if fp32 else "fp64"]
    except Exception:
        return {"bound": "fp64-pipe", "achieved": None, "peak": None, "frac": None, "unit": "TFLOP/s", "traffic": None,
                "note": "profiles/r02_rollout_ops.json missing: run scripts/rollout_ops.py under ncu"}
    d = ops["dadd"] + ops["dmul"] + 2 * ops["dfma"]
    f = ops["fadd"] + ops["fmul"] + 2 * ops["ffma"]
    peak64 = n_sms * 64 * 2 * sm_mhz * 1e6 / 1e12
    peak32 = n_sms * 128 * 2 * sm_mhz * 1e6 / 1e12
    a64 = d / (kernel_ms * 1e-3) / 1e12
    a32 = f / (kernel_ms * 1e-3) / 1e12
    main64 = not fp32
    return {"bound": "fp64-pipe (CUDA cores)" if main64 else "fp32-pipe (CUDA cores)", "kernel": "rollout_kernel (all candidates + goal-biased continuations of one round)",
            "achieved": a64 if main64 else a32, "peak": peak64 if main64 else peak32, "unit": "TFLOP/s (executed: retired FP thread instructions, fma = 2)",
            "frac": (a64 / peak64) if main64 else (a32 / peak32),
            "peak_source": f"derived: {n_sms} SMs x {'64 FP64' if main64 else '128 FP32'} lanes x 2 x {sm_mhz:.0f} MHz sampled in-run (B200_PROFILING.md has no CUDA-core figure)",
            "kernel_ms_per_launch": kernel_ms,
            "executed_flop_per_launch": {"fp64": d, "fp32": f},
            "other_pipe": {"fp32_frac" if main64 else "fp64_frac": (a32 / peak32) if main64 else (a64 / peak64)},
            "executed_flop_per_sim_step": (d + f) / max(1, ops.get("sim_steps_per_launch", per_launch_steps)),
            "issue_slot_frac": ops.get("issue_active_pct", 0) / 100.0 if ops.get("issue_active_pct") else None,
            "ncu": {k: ops.get(k) for k in ("warps_active_pct", "issue_active_pct", "threads_per_inst", "pipe_fp64_pct", "pipe_fma_pct",
                                            "inst_executed", "gpu_time_ms", "source")},
            "traffic": ops.get("dram_bytes"),
            "hbm": {"algorithmic_bytes_per_launch": hbm_bytes, "achieved_gbs": hbm_bytes / (kernel_ms * 1e-3) / 1e9,
                    "peak_gbs": peaks.get("hbm_gbs"), "note": "~190 B per rollout (parent gather, sample, result); HBM is idle on this path"}}


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import clrrt_b200 as clrrt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    # nvidia-smi needs about a second before its first sample: start it before the workload is built
    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))  # control plane only: barriers, timing max
    # a dedicated (non-default) stream shared by torch and the library: torch.cuda.Event then times the library's kernels
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    prm = clrrt.default_params()
    prm.fp32 = 1 if args.fp32 else 0
    pl = clrrt.Planner(params=prm, device=local, tree_capacity=TREE_SNAPSHOT + 2 * K_ROUND * world + 1024, max_round=K_ROUND,
                       stream=stream.cuda_stream)
    if world > 1:
        # the data plane is the library's own: ncclAllGather inside clrrt_expand_round (cl-rrt_b200/csrc/exchange.cuh)
        uid = [clrrt.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        pl.comm_init(uid[0], rank, world)
    # every rank draws the global stream and keeps its contiguous shard
    boxes, smp, heu = build_workload(pl, clrrt, rank, world)
    tree_rec = load_snapshot()
    n0 = pl.tree_size()
    d_smp = torch.from_numpy(smp).cuda()
    d_heu = torch.from_numpy(heu).cuda()
    h_smp = torch.from_numpy(smp).pin_memory()
    h_heu = torch.from_numpy(heu).pin_memory()
    np_smp, np_heu = h_smp.numpy(), h_heu.numpy()  # views of the pinned buffers
    # pinned destinations of the e2e leg's result read-back: two, the copy of round r overlaps round r + 1
    h_nodes = [torch.empty(2 * K_ROUND * world * clrrt.RECORD_BYTES, dtype=torch.uint8).pin_memory() for _ in range(2)]
    np_nodes = [h.numpy().view(clrrt.NODE_DTYPE) for h in h_nodes]

    def one_round(dev_inputs=True):
        if dev_inputs:
            return pl.expand_round_dev(d_smp.data_ptr(), d_heu.data_ptr(), K_ROUND)
        # the reference-facing call: clrrt_expand_round with HOST buffers (pinned); the library copies them to the device
        # inside the timed region
        return pl.expand_round(np_smp, np_heu)

    def timed(dev_inputs, steps, download, min_seconds=0.0):
        tot_steps = tot_roll = 0
        ms_roll = ms_ex = 0.0
        launches = 0
        d2h = 0
        done = 0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        e0.record(stream)
        while done < steps or (min_seconds > 0 and world == 1 and time.perf_counter() - t_wall < min_seconds):
            st = one_round(dev_inputs)
            if download:
                # the step's result: the nodes the round appended (all ranks' nodes: every rank holds the full tree), read back
                # into pinned memory on a second stream while the next round runs; the last copy is awaited inside the timed region
                nodes = pl.tree_download_range_async(n0, pl.tree_size() - n0, np_nodes[done & 1])
                d2h = nodes.nbytes + ctypes.sizeof(clrrt.RoundStats)
            pl.tree_truncate(n0)
            tot_steps += st.sim_steps
            tot_roll += st.rollouts
            ms_roll += st.ms_rollout
            ms_ex += st.ms_exchange
            # nn_bin, nn_scan, nn_scatter, nn_tile, nearest_sorted, ref_end, order_scan, order_scatter, setup, rollout, select,
            # scan_block_sums, scan_sums, pack_records, append (append_records, or append_gathered after the all-gather);
            # the e2e leg adds export_nodes for the read-back
            launches += 15 + (1 if download else 0)
            done += 1
        if download:
            pl.download_wait()
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, float(tot_steps), float(tot_roll)], dtype=torch.float64, device="cuda")
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, tot_steps, tot_roll = float(tmax[0]), int(t[1]), int(t[2])
        return dict(ms=ms, steps=tot_steps, rollouts=tot_roll, ms_roll=ms_roll, ms_ex=ms_ex, launches=launches, d2h=d2h, st=st, rounds=done)

    for _ in range(args.warmup):
        one_round(True)
        pl.tree_truncate(n0)
    if rank == 0:
        sampler.mark_begin()
    dev = timed(True, args.steps, False)
    # the end-to-end leg gets the same W untimed warm-up steps of ITS path (host-buffer call, result read-back): their
    # first use allocates the export buffer and loads the export kernel
    timed(False, args.warmup, True)
    e2e = timed(False, args.steps, True)
    # the same device-resident step over a longer window (>= 2 s): the number a 0.1 s timed region cannot vouch for
    sus = timed(True, args.steps, False, min_seconds=args.sustain_seconds) if (world == 1 and args.sustain_seconds > 0) else None
    if rank == 0:
        sampler.mark_end()
    clocks = sampler.summary() if rank == 0 else None
    digest = None
    if world > 1:
        # the trees of all ranks after one round are the same tree
        one_round(True)
        digests = [None] * world
        dist.all_gather_object(digests, pl.tree_digest())
        digest = {"tree_digest": digests[0], "digest_equal_on_all_ranks": all(d == digests[0] for d in digests)}
        pl.tree_truncate(n0)
    else:
        one_round(True)
        digest = {"tree_digest": pl.tree_digest(), "digest_equal_on_all_ranks": True}
        pl.tree_truncate(n0)
    pl.close()
    c4 = None
    if not args.no_c4:
        c4 = c4_strong(clrrt, torch, dist, rank, world, local, stream)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    steps = args.steps
    value = dev["steps"] / (dev["ms"] * 1e-3)
    st = dev["st"]
    kernel_ms = dev["ms_roll"] / dev["rounds"]
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    props = torch.cuda.get_device_properties(local)
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {"hbm_gbs": 6650.0}
    hbm_bytes = 190.0 * (dev["rollouts"] / max(1, dev["rounds"]) / world)
    roof = executed_roofline(kernel_ms, sm_mhz, props.multi_processor_count, args.fp32, peaks, st.sim_steps, hbm_bytes)
    # the reference-algorithm view (what round 1 reported as frac): how much arithmetic the O(1) waypoint cursor and the
    # pose grid save over the reference's formulation — a statement about the algorithm, not about pipe occupancy
    cpu = None
    R_A = work_profile(boxes, tree_rec, smp, heu)
    flop_step = algorithmic_flop_per_step(len(boxes), *R_A)
    roof["algorithmic_saving"] = {
        "reference_flop_per_sim_step": flop_step, "points_scanned_per_step_R": R_A[0], "sat_axes_per_pair_A": R_A[1],
        "reference_formulation_tflops_equivalent": flop_step * st.sim_steps / (kernel_ms * 1e-3) / 1e12,
        "note": "SURVEY.md §8d count of the REFERENCE algorithm (O(N) waypoint scan, SAT against every obstacle) for the steps executed; "
                "not a roofline: the kernel does not execute this work"}
    if not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(boxes, tree_rec, smp, heu, budget_s=args.cpu_seconds)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dev["ms"] / dev["rounds"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ("f32 rollout (tolerance mode) + f32 SAT/Dubins" if args.fp32 else "f64 rollout + f32 SAT/Dubins (the reference's mixture)"),
        "data": "synthetic", "config": workload_config(world, args.fp32),
        "rollouts_per_s": dev["rollouts"] / (dev["ms"] * 1e-3), "sim_steps_per_round": dev["steps"] / dev["rounds"],
        "e2e": {"value": e2e["steps"] / (e2e["ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(smp.nbytes + heu.nbytes),
                "d2h_bytes_per_step": int(e2e["d2h"]), "ms_per_step": e2e["ms"] / e2e["rounds"]},
        "gpu_launches": dev["launches"], "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "exchange_ms_per_round": dev["ms_ex"] / dev["rounds"] if world > 1 else 0.0,
        "tree_after_one_round": digest,
    }
    if sus:
        line["sustained"] = {"value": sus["steps"] / (sus["ms"] * 1e-3), "unit": UNIT, "seconds": sus["ms"] * 1e-3, "rounds": sus["rounds"],
                             "ms_per_step": sus["ms"] / sus["rounds"]}
    if c4:
        line["c4"] = c4
    if world == 1 and not args.quick:
        extras = (("query_200ms", lambda: query_200ms_rounds(clrrt, local)), ("query_200ms_k1", lambda: query_200ms_k1(clrrt, local)),
                  ("c2", lambda: c2_batch(clrrt, local)), ("c5", lambda: c5_loop(clrrt, local)))
        for name, fn in extras:
            try:
                line[name] = fn()
            except Exception as e:  # noqa: BLE001 — a sub-record must not take the headline down
                line[name] = {"error": f"{type(e).__name__}: {e}"}
        if not args.fp32:
            try:
                line["fp32"] = fp32_subrecord(clrrt, torch, local, stream, boxes, tree_rec, smp, heu, props, sm_mhz, peaks)
            except Exception as e:  # noqa: BLE001
                line["fp32"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def fp32_subrecord(clrrt, torch, local, stream, boxes, tree_rec, smp, heu, props, sm_mhz, peaks, steps=10, warmup=3):
    """The same C3 round in fp32 mode (float rollout; tolerance stated and checked in tests/test_gpu_fp32.py)."""
    prm = clrrt.default_params()
    prm.fp32 = 1
    pl = clrrt.Planner(params=prm, device=local, tree_capacity=TREE_SNAPSHOT + 2 * K_ROUND + 1024, max_round=K_ROUND, stream=stream.cuda_stream)
    pl.set_query(CAR, GOAL, VMAX)
    pl.set_obstacles(boxes)
    pl.tree_reset_records(tree_rec)
    n0 = pl.tree_size()
    d_s, d_h = torch.from_numpy(smp).cuda(), torch.from_numpy(heu).cuda()
    for _ in range(warmup):
        pl.expand_round_dev(d_s.data_ptr(), d_h.data_ptr(), K_ROUND)
        pl.tree_truncate(n0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    tot = 0
    ms_roll = 0.0
    for _ in range(steps):
        st = pl.expand_round_dev(d_s.data_ptr(), d_h.data_ptr(), K_ROUND)
        tot += st.sim_steps
        ms_roll += st.ms_rollout
        pl.tree_truncate(n0)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    pl.close()
    roof = executed_roofline(ms_roll / steps, sm_mhz, props.multi_processor_count, True, peaks, st.sim_steps, 0.0)
    return {"value": tot / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "kernel_ms_per_launch": ms_roll / steps,
            "roofline": {k: roof.get(k) for k in ("bound", "achieved", "peak", "frac", "unit", "executed_flop_per_launch", "ncu")},
            "tolerance": "tests/test_gpu_fp32.py: verdicts agree for >= 99.5 % of rollouts, accepted rollouts within 5 mm / 1e-4 rad / 1e-4 relative cost"}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation, nothing of the product on the path
# ---------------------------------------------------------------------------------------------------------------------
def _ref_kind():
    from cpulib import ref_available
    return "ref" if ref_available(False) else "oracle"


def _ref_worker(args):
    """One process = one copy of the (single-threaded, global-state) reference on a disjoint shard."""
    shard, n_per_step, steps, warmup, tree, boxes, samples, heur = args
    from cpulib import CpuPlanner
    cpu = CpuPlanner(_ref_kind())
    cpu.set_obstacles(boxes)
    cpu.tree_init(CAR, GOAL, VMAX)
    cpu.tree_import(tree)
    out = []
    for s in range(warmup + steps):
        lo = (shard * (warmup + steps) + s) * n_per_step
        t0 = time.perf_counter()
        cand, _, cnt = cpu.nearest_batch(samples[lo:lo + n_per_step], heur[lo:lo + n_per_step])
        ok = cnt > 0
        o = cpu.rollout_batch(cand[ok, 0], samples[lo:lo + n_per_step][ok])
        dt = time.perf_counter() - t0
        if s >= warmup:
            out.append((dt, int(o[:, 14].sum()), int(ok.sum())))
    return out


def query_200ms_reference(kind, budget_ms=200.0):
    """The reference's own loop (Timer(200) around expandTree, rrt/src/motionplanner.cpp:39-43) on one host core."""
    from cpulib import CpuPlanner
    cpu = CpuPlanner(kind)
    cpu.set_obstacles(scene_c1_boxes())
    cpu.srand(1)
    cpu.tree_init(C1_CAR, C1_GOAL, VMAX)
    nodes, it = cpu.expand_timed(budget_ms)
    return {"nodes": int(nodes), "iterations": int(it), "budget_ms": budget_ms, "cores": 1,
            "scene": "C1: straight road, goal 50 m ahead, 10 static boxes, Prius parameters"}


def c5_reference(kind):
    """Config C5 on the reference: the recorded loop's inputs, 100 expandTree iterations per query, commit_path = true."""
    path = os.path.join(ROOT, "tests", "golden", "g5_replan.npz")
    if not os.path.exists(path) or not kind.startswith("ref"):
        return None
    from cpulib import CpuPlanner
    g = np.load(path)
    iters, nq = int(g["iters"]), int(g["world"].shape[0])
    cpu = CpuPlanner(kind)
    cpu.srand(1)
    cpu.commit_reset()
    t0 = time.perf_counter()
    nodes = 0
    for q in range(nq):
        cpu.set_obstacles(g["obstacles"][q])
        carried, tree_n, nbest, sim_steps, cost = cpu.query_commit(g["world"][q], g["goal"][q], VMAX, iters)
        nodes += int(tree_n)
    wall = time.perf_counter() - t0
    return {"queries": nq, "iterations_per_query": iters, "wall_ms_per_query": 1e3 * wall / nq, "nodes_per_query": nodes / nq, "cores": 1}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref: the unmodified sources compiled by
    oracle/build_ref.sh; else the oracle port) on all host cores, each step a bounded sample of the same workload.  Nothing
    of the product is loaded: the tree snapshot is the committed golden, the samples come from the reference's own
    sampleAroundVehicle."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from cpulib import CpuPlanner
    kind = _ref_kind()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    boxes = scene_c3_boxes()
    tree = load_snapshot()
    cores = os.cpu_count() or 1
    n_per_step = 64
    cpu = CpuPlanner(kind)
    cpu.tree_init(CAR, GOAL, VMAX)
    cpu.srand(2)
    samples, heur, _ = cpu.draw_samples(cores * (args.steps + args.warmup) * n_per_step)
    jobs = [(w, n_per_step, args.steps, args.warmup, tree, boxes, samples, heur) for w in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    per_step = [max(r[s][0] for r in res) for s in range(args.steps)]
    steps_total = sum(x[1] for r in res for x in r)
    roll_total = sum(x[2] for r in res for x in r)
    secs = sum(per_step)
    value = steps_total / secs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 rollout + f32 SAT/Dubins (the reference's mixture)", "data": "synthetic",
            "config": workload_config(world),
            "rollouts_per_s": roll_total / secs,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if kind.startswith("ref") else "port",
                             "sample": f"each step: {n_per_step} samples per core x {cores} independent single-threaded reference processes "
                                       f"(candidate search + top-1 rollout of the same C3 round: same snapshot, scene and sample stream), "
                                       f"{roll_total} rollouts / {steps_total} sim steps in total, wall {wall:.1f} s"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    for name, fn in (("query_200ms_k1", lambda: query_200ms_reference(kind)), ("c5", lambda: c5_reference(kind))):
        try:
            line[name] = fn()
        except Exception as e:  # noqa: BLE001
            line[name] = {"error": str(e)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--sustain-seconds", type=float, default=2.0, help="extra device-resident leg of at least this many seconds (N = 1)")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 strong-scaling sub-record")
    ap.add_argument("--quick", action="store_true", help="headline only: no query / c2 / c5 / fp32 sub-records")
    ap.add_argument("--fp32", action="store_true", help="fp32 rollout mode (tolerance mode; default is the fp64 parity mode)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
