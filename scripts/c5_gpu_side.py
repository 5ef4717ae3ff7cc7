"""C5 debugging, GPU side: bisects the first expandTree call of query Q whose counters differ from the reference table
written by scripts/c5_ref_side.py.  Usage (under gpurun): python scripts/c5_gpu_side.py Q"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import clrrt_b200 as clrrt
from test_gpu_replan import HostPlanner
g = np.load(os.path.join(ROOT, "tests/golden/g5_replan.npz"))
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 41
ref = np.load(os.path.join(ROOT, "variants", f"c5_ref_q{Q}.npy"))
def run(m):
    hp = HostPlanner(samples_per_round=1, commit_path=True)
    C.CDLL(None).srand(C.c_uint(1))
    for q in range(Q):
        hp.query(g["world"][q], g["goal"][q], g["obstacles"][q], 100)
    sizes, cost, cnt = hp.query(g["world"][Q], g["goal"][Q], g["obstacles"][Q], m)
    hp.close()
    return [m, int(sizes[0]), int(sizes[1]), cnt.sim_count, cnt.fail_collision, cnt.fail_acclimit, cnt.fail_iterlimit, cnt.sim_count]
lo, hi = 0, 100   # invariant: equal at lo, different at hi
r = run(100); print("m=100 ours", r, "ref", ref[100].tolist())
while hi - lo > 1:
    mid = (lo + hi) // 2
    r = run(mid)
    same = r[2:] == ref[mid][2:].tolist()
    print("m", mid, "ours", r, "ref", ref[mid].tolist(), "same" if same else "DIFF", flush=True)
    if same: lo = mid
    else: hi = mid
print("first differing iteration count:", hi)
