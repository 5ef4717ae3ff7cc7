"""Where a C5 query's wall time goes (host facade phases), for chunk / window settings of the sequential mode."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from test_gpu_replan import HostPlanner
g = np.load(os.path.join(ROOT, "tests", "golden", "g5_replan.npz"))
iters, nq = int(g["iters"]), int(g["world"].shape[0])
for chunk, window in ((16, 0), (100, 0), (100, 2), (100, 4), (100, 8), (100, 16)):
    hp = HostPlanner(samples_per_round=1, commit_path=True)
    hp.lib.clrrt_host_planner_set_sequential.argtypes = [C.c_void_p, C.c_int, C.c_int]
    hp.lib.clrrt_host_planner_timings.argtypes = [C.c_void_p, C.c_void_p]
    hp.lib.clrrt_host_planner_set_sequential(hp.h, chunk, window)
    C.CDLL(None).srand(C.c_uint(1))
    acc = np.zeros(6); ms = np.zeros(6); nwin = 0
    t0 = time.perf_counter()
    ok = 0
    for q in range(nq):
        sizes, cost, cnt = hp.query(g["world"][q], g["goal"][q], g["obstacles"][q], iters)
        nwin += hp.lib.clrrt_host_planner_timings(hp.h, ms.ctypes.data)
        acc += ms
        ok += int(sizes[:3].tolist() == g["sizes"][q].tolist() and cnt.sim_count == int(g["sim_steps"][q]))
    wall = time.perf_counter() - t0
    print(f"chunk {chunk:3d} window {window:2d}: {1e3*wall/nq:6.2f} ms/query; phases params {acc[0]/nq:.2f} obstacles {acc[1]/nq:.2f} init {acc[2]/nq:.2f} "
          f"expand {acc[3]/nq:.2f} bestpath {acc[4]/nq:.2f} messages {acc[5]/nq:.2f}; {ok}/100 equal to the golden; {nwin/nq:.1f} windows per query")
    hp.close()
