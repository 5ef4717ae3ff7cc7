#!/usr/bin/env python
"""bench.py — CL-RRT tree-expansion throughput on B200 (contract in the task statement, tier section ④).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU code (oracle/_ref)

A *step* is one expansion round: 65 536 samples (per GPU) against a fixed 4096-node tree snapshot in the dense
urban scene (config C3: 1000 oriented-box obstacles, goal 100 m ahead): candidate search, closed-loop rollouts in
candidate order until the first success, goal-biased rollouts, ordered append (+ node all-gather when N > 1).
The tree is truncated back to the snapshot after every round so that each step does the same work.
Metric: closed-loop sim steps per second (one sim step = one iteration of rrt/src/simulation.cpp:58), whole job.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "closed_loop_sim_steps_per_s"
UNIT = "sim-steps/s"
WORKLOAD = "C3 dense urban scene: 1000 OBB obstacles, 65536 samples/round/GPU, 4096-node tree snapshot, fp64 parity mode"
K_ROUND = 65536
TREE_SNAPSHOT = 4096
CAR = (0.0, 0.0, 0.0, 0.0, 3.0, 0.0)
GOAL = (100.0, 0.0, 0.0, 0.0)
VMAX = 5.0


SCALING = "weak"
SCENE = "c3"


def select_workload(name, world):
    """--workload c4: 2^20 samples per round IN TOTAL, sharded over the ranks (strong scaling),
    the 10 boxes of C1/C2, goal 50 m ahead, 4096-node tree snapshot.  Default c3: 65 536 samples per round and GPU."""
    global WORKLOAD, K_ROUND, CAR, GOAL, SCALING, SCENE
    if name == "c4":
        K_ROUND = (1 << 20) // world
        CAR, GOAL, SCALING, SCENE = (0.0, 0.0, 0.0, 0.0, 3.0, 0.0), (50.0, 0.0, 0.0, 0.0), "strong", "c4"
        WORKLOAD = (f"C4 multi-GPU sweep: 2^20 samples/round in total ({K_ROUND} per GPU), 10 OBB obstacles, "
                    "4096-node tree snapshot, fp64 parity mode")


def scene_boxes():
    return scene_c3_boxes() if SCENE == "c3" else scene_c1_boxes()


def scene_c3_boxes():
    """closed-form layout, no RNG."""
    o = np.zeros((1000, 7))
    for i in range(1000):
        c, r = i % 100, i // 100
        y = 3.0 + 1.5 * (r // 2)
        o[i] = [5.0 + c, y if r % 2 == 0 else -y, (0.1 * i) % np.pi, 2.0, 4.0, 0.0, 0.0]
    return o


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


def build_workload(pl, clrrt, rank, world):
    """Scene, tree snapshot (grown on the GPU with this library, deterministic) and the per-rank sample shards."""
    boxes = scene_boxes()
    pl.set_query(CAR, GOAL, VMAX)
    pl.set_obstacles(boxes)
    pl.tree_reset(clrrt.root_node(CAR))
    s, h = clrrt.draw_samples(GOAL, 8192 * 12, seed=1)
    i = 0
    while pl.tree_size() < TREE_SNAPSHOT and i < 12:
        pl.expand_round(s[i * 8192:(i + 1) * 8192], h[i * 8192:(i + 1) * 8192])
        i += 1
    if pl.tree_size() < TREE_SNAPSHOT:
        raise RuntimeError(f"tree snapshot only reached {pl.tree_size()} nodes")
    pl.tree_truncate(TREE_SNAPSHOT)
    # every rank draws the global stream and keeps its contiguous shard
    gs, gh = clrrt.draw_samples(GOAL, K_ROUND * world, seed=2)
    return boxes, gs[rank * K_ROUND:(rank + 1) * K_ROUND].copy(), gh[rank * K_ROUND:(rank + 1) * K_ROUND].copy()


C1_CAR = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
C1_GOAL = (50.0, 0.0, 0.0, 0.0)


def scene_c1_boxes():
    """10 static boxes, centre (8 + 4.7 i, +3 even / -3 odd), size_x 4, size_y 8."""
    o = np.zeros((10, 7))
    for i in range(10):
        o[i] = [8 + 4.7 * i, 3.0 if i % 2 == 0 else -3.0, 0.0, 4.0, 8.0, 0.0, 0.0]
    return o


def query_200ms_ours(clrrt, device, K=16384, budget_ms=200.0):
    """Second half of BASELINE.json's metric: tree nodes grown by one planMotion query with a 200 ms expansion budget
    (config C1).  Samples are drawn on the host with the reference's expressions (rand() after srand(1)), K per round;
    the wall clock covers drawing, the host->device copy and the round."""
    pl = clrrt.Planner(device=device, tree_capacity=(1 << 20) + 2 * K, max_round=K)  # K: 1024 -> 50 k nodes, 4096 -> 113 k, 16384 -> 225 k, 65536 -> 305 k (8 rounds)
    pl.set_query(C1_CAR, C1_GOAL, VMAX)
    pl.set_obstacles(scene_c1_boxes())
    pl.tree_reset(clrrt.root_node(C1_CAR))
    s, h = clrrt.draw_samples(C1_GOAL, K, seed=1)
    pl.expand_round(s, h)  # warm-up (first launch, lazy module load), then start over
    pl.tree_reset(clrrt.root_node(C1_CAR))
    clrrt.draw_samples(C1_GOAL, 1, seed=1)
    rounds = steps = rollouts = 0
    # the next round's samples are drawn (host rand(), ~0.9 ms for 16384) while the device expands the current round: the
    # draws do not depend on the tree and stay in stream order (one drawing thread, joined before the next draw starts)
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=1)
    t0 = time.perf_counter()
    nxt = pool.submit(clrrt.draw_samples, C1_GOAL, K)
    while (time.perf_counter() - t0) * 1e3 < budget_ms and pl.tree_size() < (1 << 20):
        s, h = nxt.result()
        nxt = pool.submit(clrrt.draw_samples, C1_GOAL, K)
        st = pl.expand_round(s, h)
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
            "wall_ms": wall, "budget_ms": budget_ms, "best_path_nodes": int(path),
            "scene": "C1: straight road, goal 50 m ahead, 10 static boxes, Prius parameters"}


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


def cpu_baseline_sample(boxes, tree_records, samples, heur, budget_s=12.0, kind=None):
    """The reference's own CPU code (oracle/_ref when present, else the C restatement) on a bounded sample of the
    same workload: top-1 candidate rollouts, batches of 32 until the time budget is spent."""
    from cpulib import CpuPlanner, ref_available
    kind = kind or ("ref_defined" if ref_available(True) else "oracle")
    cpu = CpuPlanner(kind)
    cpu.set_obstacles(boxes)
    cpu.tree_init(CAR, GOAL, VMAX)
    cpu.tree_import(tree_records)
    steps = rollouts = 0
    secs = 0.0
    t_nn = 0.0
    j = 0
    R_scan = A_axes = None
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
                      f"of the same round, {secs + t_nn:.1f} s on 1 host core ({'oracle/_ref' if kind.startswith('ref') else 'oracle C port'}, -O3 -DNDEBUG, ROS logging compiled out)",
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a dedicated (non-default) stream shared by torch and the library: torch.cuda.Event then times the library's kernels
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    prm = clrrt.default_params()
    prm.fp32 = 1 if args.fp32 else 0
    pl = clrrt.Planner(params=prm, device=local, tree_capacity=TREE_SNAPSHOT + 2 * K_ROUND * world + 1024, max_round=K_ROUND,
                       stream=stream.cuda_stream)
    boxes, smp, heu = build_workload(pl, clrrt, rank, world)
    n0 = pl.tree_size()
    d_smp = torch.from_numpy(smp).cuda()
    d_heu = torch.from_numpy(heu).cuda()
    h_smp = torch.from_numpy(smp).pin_memory()
    h_heu = torch.from_numpy(heu).pin_memory()
    np_smp, np_heu = h_smp.numpy(), h_heu.numpy()  # views of the pinned buffers
    h_nodes = torch.empty(2 * K_ROUND * clrrt.RECORD_BYTES, dtype=torch.uint8).pin_memory()
    np_nodes = h_nodes.numpy().view(clrrt.NODE_DTYPE)  # pinned destination of the e2e leg's result read-back
    if world > 1:
        from clrrt_b200.exchange import gather_records
        pl.set_defer_append(True)
        rec_bytes = clrrt.RECORD_BYTES
        max_rec = 2 * K_ROUND
        gathered = torch.empty(world * max_rec * rec_bytes, dtype=torch.uint8, device="cuda")
        counts_t = torch.zeros(world, dtype=torch.int32, device="cuda")
        h_rec = torch.empty(max_rec * rec_bytes, dtype=torch.uint8).pin_memory()

    def exchange():
        """per-round node all-gather over NCCL (cl-rrt_b200/exchange.py): per-rank counts, then fixed-stride records;
        every rank appends all ranks' chunks in rank order (= global sample order), so the trees stay identical."""
        ptr, n = pl.round_records()
        src = _as_cuda_tensor(ptr, max_rec * rec_bytes, local)
        out, counts, stride = gather_records(src, n, world, counts_t, gathered, sync=False)  # planner and torch share `stream`
        if stride == 0:
            return 0
        pl.append_records(out.data_ptr(), counts, stride)
        return int(counts.sum())

    def one_round(dev_inputs=True):
        if dev_inputs:
            st = pl.expand_round_dev(d_smp.data_ptr(), d_heu.data_ptr(), K_ROUND)
        else:
            # the reference-facing call: clrrt_expand_round with HOST buffers (pinned); the library copies them to the
            # device inside the timed region
            st = pl.expand_round(np_smp, np_heu)
        added = exchange() if world > 1 else st.nodes_added
        return st, added

    def timed(dev_inputs, steps, download):
        tot_steps = tot_roll = 0
        ms_roll = 0.0
        launches = 0
        d2h = 0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            st, added = one_round(dev_inputs)
            if download:  # the step's result: the accepted node records, back on the host
                if world == 1:
                    nodes = pl.tree_download_range(n0, pl.tree_size() - n0, out=np_nodes)
                    d2h = nodes.nbytes + ctypes.sizeof(clrrt.RoundStats)
                else:
                    # every rank reads back the records of ITS shard of the round (their union over ranks is the result)
                    ptr, n_loc = pl.round_records()
                    nb = n_loc * rec_bytes
                    if nb:
                        h_rec[:nb].copy_(_as_cuda_tensor(ptr, max_rec * rec_bytes, local)[:nb])
                    d2h = nb + ctypes.sizeof(clrrt.RoundStats)
            pl.tree_truncate(n0)
            tot_steps += st.sim_steps
            tot_roll += st.rollouts
            ms_roll += st.ms_rollout
            # nn_bin, nn_scan, nn_scatter, nn_tile, nearest_sorted, ref_end, order_scan, order_scatter, setup, rollout, select,
            # scan_block_sums, scan_sums, pack_records, append_records (single GPU: appended inside the round; multi GPU:
            # one append per rank chunk); the e2e leg adds export_nodes for the read-back
            launches += 15 if world == 1 else 14 + world
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
        return ms, tot_steps, tot_roll, ms_roll, launches, d2h, st

    for _ in range(args.warmup):
        one_round(True)
        pl.tree_truncate(n0)
    if rank == 0:
        sampler.mark_begin()
    ms, tot_steps, tot_roll, ms_roll, launches, _, st = timed(True, args.steps, False)
    # the end-to-end leg gets the same W untimed warm-up steps of ITS path (host-buffer call, result read-back): their
    # first use allocates the export buffer and loads the export kernel
    timed(False, args.warmup, True)
    e_ms, e_steps, e_roll, _, _, d2h, _ = timed(False, args.steps, True)
    if rank == 0:
        sampler.mark_end()
    clocks = sampler.summary() if rank == 0 else None
    if rank != 0:
        pl.close()
        if world > 1:
            dist.destroy_process_group()
        return
    value = tot_steps / (ms * 1e-3)
    # roofline of the dominant kernel (rollout_kernel<false>), measured live with the library's CUDA events on the
    # launch stream: algorithmic FLOP of the steps it executed / its device time, against the FP32 pipe peak at the
    # SM clock sampled during the run.
    cpu = None
    tree_rec = pl.tree_download_records()[:n0]
    R_A = work_profile(boxes, tree_rec, smp, heu)
    if not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(boxes, tree_rec, smp, heu, budget_s=args.cpu_seconds)
    n_obs = len(boxes)
    flop_step = algorithmic_flop_per_step(n_obs, *R_A)
    per_launch_steps = st.sim_steps  # last round, this rank
    kernel_ms = ms_roll / args.steps
    achieved = flop_step * per_launch_steps / (kernel_ms * 1e-3) / 1e12
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    props = torch.cuda.get_device_properties(local)
    peak = props.multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
    hbm_bytes = 190.0 * (tot_roll / max(1, args.steps) / world)
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {"hbm_gbs": 6650.0}
    # figures that only a profiler can give (DRAM traffic, pipe utilisation of the executed instructions): from the committed
    # ncu capture of this same command (profiles/, written by scripts/summarize_ncu.py); absent -> null
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "rollout_kernel_ncu_latest.json")))
    except Exception:
        ncu = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": SCALING, "vs_baseline": None,
        "dtype": ("f32 rollout (tolerance mode) + f32 SAT/Dubins" if args.fp32 else "f64 rollout + f32 SAT/Dubins (the reference's mixture)"),
        "data": "synthetic",
        "config": {"workload": WORKLOAD if not args.fp32 else WORKLOAD.replace("fp64 parity mode", "fp32 mode"), "samples_per_round_per_gpu": K_ROUND, "tree_nodes": n0, "obstacles": n_obs,
                   "parallelism": f"samples sharded over {world} GPU(s), tree replicated, node all-gather per round" if world > 1 else "1 GPU",
                   "l2_policy": "per-round inputs+outputs (samples, candidate lists, staging SoA, records: ~40 MB) are rewritten every round; the kernel is compute-bound, no L2 flush needed"},
        "rollouts_per_s": tot_roll / (ms * 1e-3), "sim_steps_per_round": tot_steps / args.steps,
        "e2e": {"value": e_steps / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(smp.nbytes + heu.nbytes),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "fp32-pipe (CUDA cores; no tensor work on this path)", "kernel": "rollout_kernel (candidates + goal-biased continuations of one round)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s (algorithmic, SURVEY.md §8d count)",
                     "frac": achieved / peak,
                     "peak_source": f"derived: {props.multi_processor_count} SMs x 128 FP32 lanes x 2 x {sm_mhz:.0f} MHz sampled in-run",
                     "kernel_ms_per_launch": kernel_ms, "algorithmic_flop_per_sim_step": flop_step,
                     "reference_work_profile": {"points_scanned_per_step_R": R_A[0], "sat_axes_per_pair_A": R_A[1]},
                     "traffic": (ncu or {}).get("dram_bytes_per_launch"),
                     "note": "achieved counts the REFERENCE algorithm's arithmetic (O(N) waypoint scan, SAT against every obstacle) for the "
                             "steps executed, so frac > 1 measures the algorithmic saving; the executed-instruction view is `ncu`",
                     "ncu": ncu,
                     "hbm": {"achieved_gbs": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                             "note": "algorithmic ~190 B per rollout; HBM is idle on this path"}},
        "cpu_baseline": cpu,
    }
    pl.close()
    if world == 1:
        line["query_200ms"] = query_200ms_ours(clrrt, local)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _as_cuda_tensor(ptr, nbytes, device):
    """Zero-copy torch view of a device buffer owned by the library (for NCCL all-gather)."""
    import torch

    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    return torch.as_tensor(h, device=torch.device("cuda", device))


def _ref_worker(args):
    """One process = one copy of the (single-threaded, global-state) reference on a disjoint shard."""
    shard, n_per_step, steps, warmup, tree, boxes, samples, heur = args
    from cpulib import CpuPlanner, ref_available
    cpu = CpuPlanner("ref_defined" if ref_available(True) else "oracle")
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


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref, else the oracle port) on all host
    cores, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from cpulib import CpuPlanner, ref_available
    kind = "ref_defined" if ref_available(True) else "oracle"
    boxes = scene_boxes()
    # the same tree snapshot needs the GPU library; without a GPU grow a smaller one with the CPU code itself
    tree = None
    try:
        import torch
        if torch.cuda.is_available():
            import clrrt_b200 as clrrt
            pl = clrrt.Planner(device=0, tree_capacity=TREE_SNAPSHOT + 2 * K_ROUND + 1024, max_round=K_ROUND)
            build_workload(pl, clrrt, 0, 1)
            tree = pl.tree_download_records()[:TREE_SNAPSHOT]
            pl.close()
    except Exception:
        tree = None
    if tree is None:
        cpu = CpuPlanner(kind)
        cpu.set_obstacles(boxes)
        cpu.srand(1)
        cpu.tree_init(CAR, GOAL, VMAX)
        cpu.expand(300)
        tree = cpu.tree_export()
    import clrrt_b200 as clrrt
    cores = os.cpu_count() or 1
    n_per_step = 64
    samples, heur = clrrt.draw_samples(GOAL, cores * (args.steps + args.warmup) * n_per_step, seed=2)
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
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 rollout + f32 SAT/Dubins", "data": "synthetic",
            "config": {"workload": WORKLOAD, "tree_nodes": int(len(tree)), "obstacles": int(len(boxes))},
            "rollouts_per_s": roll_total / secs,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if kind.startswith("ref") else "port",
                             "sample": f"each step: {n_per_step} samples per core x {cores} independent single-threaded reference processes "
                                       f"(candidate search + top-1 rollout), {roll_total} rollouts / {steps_total} sim steps in total, wall {wall:.1f} s"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    try:
        line["query_200ms"] = query_200ms_reference(kind)
    except Exception as e:  # noqa: BLE001
        line["query_200ms"] = {"error": str(e)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--workload", default="c3", choices=["c3", "c4"], help="c3 (default): the headline single-GPU configuration, weak "
                    "scaling; c4: 2^20 samples per round in total, strong scaling")
    ap.add_argument("--fp32", action="store_true", help="fp32 rollout mode (tolerance mode; default is the fp64 parity mode)")
    args = ap.parse_args()
    select_workload(args.workload, int(os.environ.get("WORLD_SIZE", "1")))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
