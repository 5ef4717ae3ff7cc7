"""Per-phase clock accounting of the rollout kernels (needs a -DCLRRT_PHASE_CLOCKS build selected with CLRRT_LIB)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import clrrt_b200 as clrrt
import bench
pl = clrrt.Planner(device=0, tree_capacity=bench.TREE_SNAPSHOT + 2 * bench.K_ROUND + 1024, max_round=bench.K_ROUND)
lib = clrrt.load_library()
lib.clrrt_debug_phase_clocks.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
def clocks(reset=True):
    out = (C.c_ulonglong * 16)()
    lib.clrrt_debug_phase_clocks(pl.h, out, 1 if reset else 0)
    return np.array(list(out), dtype=np.float64)
def report(tag):
    c = clocks()
    ws = max(1.0, c[4])
    print(f"{tag}: warp-steps {int(c[4])}; cycles per warp-step: refill {c[0]/ws:.0f} dynamics {c[1]/ws:.0f} collision [lookup {c[5]/ws:.0f} coarse {c[6]/ws:.0f} fine {c[7]/ws:.0f} narrow+rest {c[2]/ws:.0f}] finish {c[3]/ws:.0f} total {(c[0]+c[1]+c[2]+c[3]+c[5]+c[6]+c[7])/ws:.0f}; active lanes/warp-step {c[8]/ws:.1f}; after the queue drained: {int(c[9])} warp-steps ({100*c[9]/ws:.0f}%) at {c[10]/max(1,c[9]):.1f} lanes")
boxes, smp, heu = bench.build_workload(pl, clrrt, 0, 1)
n0 = pl.tree_size(); clocks()
pl.propagate_batch([0], [[60.0, 0.0]], [0]); report("lone rollout, corridor centre")
pl.propagate_batch([0] * 32, [[60.0, 0.1 * i - 1.5] for i in range(32)], [0] * 32); report("one full warp")
pl.propagate_batch([0], [[0.0, 0.0]], [1]); report("lone goal-biased rollout")
for K in (4096, 65536):
    st = pl.expand_round(smp[:K], heu[:K]); pl.tree_truncate(n0); report(f"round K={K} (main+gb)")
