"""Per-source-line stall samples of a kernel from `ncu -i X.ncu-rep --page source --csv` output (diagnostic)."""
import csv, sys, collections
csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2] if len(sys.argv) > 2 else "rollout_kernel"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
cur_file = cur_fn = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": continue
    if pat not in (cur_fn or "") or r[2] != "-": continue
    try: ln = int(r[0]); smp = int(r[4] or 0); inst = int(r[7] or 0); thr = int(r[8] or 0)
    except Exception: continue
    k = (cur_file, ln); agg[k][0] += smp; agg[k][1] += inst; agg[k][2] += thr; agg[k][3] = r[1][:100]
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print("samples", ts, "warp instructions", ti)
byf = collections.defaultdict(lambda: [0, 0])
for (f, l), v in agg.items(): byf[f][0] += v[0]; byf[f][1] += v[1]
for f, v in byf.items(): print(f"  {f}: samples {100*v[0]/max(1,ts):.1f}% inst {100*v[1]/max(1,ti):.1f}%")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{l} S{100*v[0]/max(1,ts):.2f}% I{100*v[1]/max(1,ti):.2f}% T{v[2]/max(1,v[1]):.1f} {v[3]}")
