"""Nodes grown by one 200 ms query (config C1) against the samples-per-round setting (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import clrrt_b200 as clrrt
import bench
for K in (1024, 4096, 16384, 65536):
    r = bench.query_200ms_rounds(clrrt, 0, K=K)
    print(K, {k: r[k] for k in ("nodes", "rounds", "sim_steps", "wall_ms", "best_path_nodes")})
