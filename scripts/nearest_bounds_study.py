"""Offline study (numpy) of tile-level bounds for the optimise key on a large C1 tree: how many tiles the search must
visit under (a) the bound in use (0.999 dist(sample, tile box) + min costE), (b) a projected bound
min(costE - 0.999 node.e_k) + 0.999 sample.e_k over a fan of directions e_k, (c) the same with the nodes of every spatial
bin ordered by excess cost (costE - |node - root|).  Diagnostic."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import clrrt_b200 as clrrt, bench
K = 16384
pl = clrrt.Planner(device=0, tree_capacity=(1 << 20) + 2 * K, max_round=K)
pl.set_query(bench.C1_CAR, bench.C1_GOAL, bench.VMAX); pl.set_obstacles(bench.scene_c1_boxes())
pl.tree_reset(clrrt.root_node(bench.C1_CAR))
clrrt.draw_samples(bench.C1_GOAL, 1, seed=1)
for r in range(int(os.environ.get("ROUNDS", "24"))):
    s, h = clrrt.draw_samples(bench.C1_GOAL, K); pl.expand_round(s, h)
tree = pl.tree_download(); n = len(tree)
x, y = tree["state"][:, 0].astype(np.float32), tree["state"][:, 1].astype(np.float32); ce = tree["costE"]
gx, gy = bench.C1_GOAL[0], bench.C1_GOAL[1]
beta = np.arctan2(gy, gx); cb, sb = np.cos(beta), np.sin(beta); dgoal = np.hypot(gx, gy)
u = x * cb + y * sb; v = y * cb - x * sb
nl = 16 if n >= 98304 else 4 if n >= 24576 else 1
ns = 1024 // nl
bw = (dgoal + 40.0) / ns; vw = 16.0 / nl
slab = np.clip(((u + 15.0) / bw).astype(int), 0, ns - 1); lat = np.clip(((v + 8.0) / vw).astype(int), 0, nl - 1)
b = slab * nl + lat
excess = ce - np.hypot(x - x[0], y - y[0])
print("tree", n, "excess percentiles", np.percentile(excess, [0, 1, 5, 25, 50, 75, 95]))
S, H = clrrt.draw_samples(bench.C1_GOAL, 48)
cand, key, cnt = pl.nearest_batch(S, np.ones_like(H))
dirs = np.radians(np.arange(-40, 41, 10.0))
def study(order, name):
    tiles = [order[i:i + 256] for i in range(0, n, 256)]
    tb = []
    for t in tiles:
        tb.append((u[t].min(), u[t].max(), v[t].min(), v[t].max(), ce[t].min(),
                   [np.min(ce[t] - 0.999 * (u[t] * np.cos(d) + v[t] * np.sin(d))) for d in dirs]))
    res = []
    for j in range(len(S)):
        T = key[j, 9] if cnt[j] >= 10 else np.inf
        su = S[j, 0] * cb + S[j, 1] * sb; sv = S[j, 1] * cb - S[j, 0] * sb
        wa = wb = 0
        for (ulo, uhi, vlo, vhi, cm, g) in tb:
            du = max(ulo - su, su - uhi, 0); dv = max(vlo - sv, sv - vhi, 0)
            a_ok = 0.999 * np.hypot(du, dv) + cm <= T
            pb = max(gk + 0.999 * (su * np.cos(d) + sv * np.sin(d)) for gk, d in zip(g, dirs))
            wa += a_ok; wb += a_ok and pb <= T
        need = np.sum(0.999 * np.hypot(u - su, v - sv) + ce <= T)
        res.append((wa, wb, need, T - np.hypot(S[j, 0] - x[0], S[j, 1] - y[0])))
    res = np.array(res)
    print(f"{name}: tiles {len(tiles)}; wanted by (a) mean {res[:,0].mean():.0f}, by (a)+(b) mean {res[:,1].mean():.0f}; nodes within the node-level bound mean {res[:,2].mean():.0f}; T - |root->s| median {np.median(res[:,3]):.3f}")
study(np.argsort(b, kind="stable"), "spatial bins")
study(np.lexsort((excess, b)), "spatial bins, by excess inside a bin")
cls = np.digitize(excess, [0.1, 0.5, 2.0])
nl2 = max(nl // 4, 1); ns2 = ns
lat2 = np.clip(((v + 8.0) / (16.0 / nl2)).astype(int), 0, nl2 - 1)
study(np.lexsort((cls, slab * nl2 + lat2)), "coarser lateral bins x 4 excess classes")
