"""Offline study (numpy) of a tile-level FEASIBILITY bound for the explore key on a large C1 tree, under several node
orders: how many tiles pass the box bound, and how many of those also pass the direction-class feasibility bound
(no class of the tile has S . f_c >= min rb . f_c).  Diagnostic; conclusions in profiles/r02_experiments.md."""
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
x, y = tree["state"][:, 0], tree["state"][:, 1]
rb = tree["ref_back"]; rf = tree["ref_front"]
gx, gy = bench.C1_GOAL[0], bench.C1_GOAL[1]
beta = np.arctan2(gy, gx); cb, sb = np.cos(beta), np.sin(beta); dgoal = np.hypot(gx, gy)
rot = lambda px, py: (px * cb + py * sb, py * cb - px * sb)
u, v = rot(x, y); ru, rv = rot(rb[:, 0], rb[:, 1]); du_, dv_ = rot(rb[:, 0] - rf[:, 0], rb[:, 1] - rf[:, 1])
th = np.arctan2(dv_, du_)
NC = 16
cls = np.clip(((th + np.pi) * NC / (2 * np.pi)).astype(int), 0, NC - 1)
fa = -np.pi + (np.arange(NC) + 0.5) * 2 * np.pi / NC
S, H = clrrt.draw_samples(bench.C1_GOAL, 64)
cand, key, cnt = pl.nearest_batch(S, np.zeros_like(H))
def bins(nslab, nlat):
    bw = (dgoal + 40.0) / nslab; vw = 16.0 / nlat
    return np.clip(((u + 15.0) / bw).astype(int), 0, nslab - 1), np.clip(((v + 8.0) / vw).astype(int), 0, nlat - 1)
def study(order, name):
    nt = (n + 255) // 256
    pad = nt * 256 - n
    o = np.concatenate([order, np.full(pad, order[-1])]).reshape(nt, 256)
    ulo, uhi, vlo, vhi = u[o].min(1), u[o].max(1), v[o].min(1), v[o].max(1)
    proj = ru[o][:, :, None] * np.cos(fa)[None, None, :] + rv[o][:, :, None] * np.sin(fa)[None, None, :]
    mask = cls[o][:, :, None] == np.arange(NC)[None, None, :]
    feas = np.where(mask, proj, np.inf).min(1)     # [nt][NC]
    res = []
    for j in range(len(S)):
        T = key[j, 9] if cnt[j] >= 10 else np.inf
        su, sv = rot(S[j, 0], S[j, 1])
        d = np.hypot(np.maximum(np.maximum(ulo - su, su - uhi), 0), np.maximum(np.maximum(vlo - sv, sv - vhi), 0))
        box = 0.999 * d <= T
        sp = su * np.cos(fa) + sv * np.sin(fa)
        ok = (sp[None, :] >= feas).any(1)
        # exact: tiles that hold a node that survives stage 1 and is feasible
        w_u, w_v = su - ru, sv - rv
        dot = w_u * du_ + w_v * dv_; crs = w_u * dv_ - w_v * du_
        fe = (dot >= np.abs(crs)) & (w_u ** 2 + w_v ** 2 >= 0.42 ** 2) & (0.999 * np.hypot(u - su, v - sv) <= T)
        need = fe[o].any(1)
        res.append((box.sum(), (box & ok).sum(), need.sum()))
    res = np.array(res, float)
    print(f"{name}: box bound {res[:,0].mean():.0f} tiles/sample, box + feasibility bound {res[:,1].mean():.0f}, tiles that hold a feasible node within T {res[:,2].mean():.0f}")
sl, la = bins(64, 16)
study(np.argsort(sl * 16 + la, kind="stable"), "64 slabs x 16 lateral (in use)")
c4 = np.clip(((th + np.pi / 2) / (np.pi / 4)).astype(int), -1, 4) + 1   # 6 classes: < -90, 4 x 45 degrees, > 90
sl, la = bins(64, 4)
study(np.lexsort((c4, la, sl)), "64 slabs x 4 lateral x 6 direction classes")
sl, la = bins(32, 8)
study(np.lexsort((c4, la, sl)), "32 slabs x 8 lateral x 6 direction classes")
sl, la = bins(64, 16)
study(np.lexsort((cls, la, sl)), "64 slabs x 16 lateral, by direction class inside a bin")
sl, la = bins(64, 1)
study(np.lexsort((cls, sl)), "64 slabs x 16 direction classes")
