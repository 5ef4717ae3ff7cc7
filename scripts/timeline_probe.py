"""Timeline of one C3 round's rollouts (needs a -DCLRRT_PHASE_CLOCKS build selected with CLRRT_LIB): when candidates and
goal-biased continuations start and end, which chains finish last, per-step latency by phase of the launch."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import clrrt_b200 as clrrt
import bench
K = int(os.environ.get("K", bench.K_ROUND))
pl = clrrt.Planner(device=0, tree_capacity=bench.TREE_SNAPSHOT + 2 * bench.K_ROUND + 1024, max_round=bench.K_ROUND)
lib = clrrt.load_library()
lib.clrrt_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
boxes, smp, heu = bench.build_workload(pl, clrrt, 0, 1)
n0 = pl.tree_size()
for rep in range(2):
    st = pl.expand_round(smp[:K], heu[:K]); pl.tree_truncate(n0)
tl = np.zeros((K * 11, 3), dtype=np.uint64)
rc = lib.clrrt_debug_timeline(pl.h, tl.ctypes.data, K)
assert rc == 0, rc
ran = tl[:, 1] > 0
t0 = tl[ran, 0].min()
start = (tl[:, 0].astype(np.float64) - float(t0)) * 1e-6
end = (tl[:, 1].astype(np.float64) - float(t0)) * 1e-6
steps = (tl[:, 2] & np.uint64(0xffff)).astype(np.int64)
code = ((tl[:, 2] >> np.uint64(16)) & np.uint64(0xff)).astype(np.int64)
gb = np.zeros(K * 11, bool); gb[K * 10:] = True
rank = np.concatenate([np.tile(np.arange(10), K), np.zeros(K, int)])
print(f"K={K}: rollout kernel {st.ms_rollout:.2f} ms; {ran.sum()} rollouts recorded ({(ran & gb).sum()} goal-biased), last end {end[ran].max():.2f} ms")
for name, m in (("candidates", ran & ~gb), ("goal-biased", ran & gb)):
    s = steps[m]
    print(f"  {name}: n {m.sum()} steps mean {s.mean():.1f} p50 {np.percentile(s,50):.0f} p90 {np.percentile(s,90):.0f} p99 {np.percentile(s,99):.0f} max {s.max()}; total lane-steps {s.sum()}; codes {np.bincount(code[m], minlength=10)}")
    print(f"    start time percentiles (ms) p50 {np.percentile(start[m],50):.2f} p90 {np.percentile(start[m],90):.2f} p99 {np.percentile(start[m],99):.2f} max {start[m].max():.2f}")
print("  lane-steps in flight by time (end-time histogram, 0.25 ms bins): work finishing in each bin / its rollouts")
edges = np.arange(0, end[ran].max() + 0.25, 0.25)
for lo in edges:
    m = ran & (end >= lo) & (end < lo + 0.25)
    if m.sum():
        d = (end[m] - start[m]) * 1e3 / np.maximum(1, steps[m])
        print(f"    [{lo:4.2f},{lo+0.25:4.2f}) rollouts {m.sum():6d} gb {int((m & gb).sum()):5d} steps {steps[m].sum():8d} long(>=200) {int((steps[m] >= 200).sum()):5d} us/step of long ones {np.mean(d[steps[m] >= 200]) if (steps[m] >= 200).any() else 0:.2f}")
# the chains that end last
order = np.argsort(-end * ran)[:12]
print("  last finishers: slot kind rank start end steps code us/step | their parent's candidate (for goal-biased): start end steps rank")
word_rank = {}
for o in order:
    if gb[o]:
        j = o - 10 * K
        c = [10 * j + r for r in range(10) if ran[10 * j + r] and code[10 * j + r] in (4, 5)]
        w = min(c) if c else -1
        extra = f"winner rank {w - 10*j} start {start[w]:.2f} end {end[w]:.2f} steps {steps[w]}; all cand ends {[round(float(end[10*j+r]),2) for r in range(10) if ran[10*j+r]]}" if w >= 0 else ""
    else:
        extra = ""
    print(f"    {o:7d} {'gb' if gb[o] else 'cand'} r{rank[o]} {start[o]:.2f} {end[o]:.2f} {steps[o]:4d} c{code[o]} {(end[o]-start[o])*1e3/max(1,steps[o]):.2f} | {extra}")
# how late do long goal-biased continuations start?
m = ran & gb & (steps >= 200)
print(f"  long goal-biased (>=200 steps): n {m.sum()} start p50 {np.percentile(start[m],50):.2f} p90 {np.percentile(start[m],90):.2f} max {start[m].max():.2f}; end p50 {np.percentile(end[m],50):.2f} p90 {np.percentile(end[m],90):.2f}")
m2 = ran & ~gb & (steps >= 200)
print(f"  long candidates (>=200 steps): n {m2.sum()} by rank {np.bincount(rank[m2], minlength=10)} start p50 {np.percentile(start[m2],50):.2f} p90 {np.percentile(start[m2],90):.2f} max {start[m2].max():.2f}")
np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"timeline_K{K}.npz"), tl=tl, smp=smp[:K])
