"""Instruction-cache footprint of a kernel from an ncu source page (ncu -i X.ncu-rep --page source --csv > src.csv):
how many 128-byte lines are executed in at least a given share of the warp steps (diagnostic)."""
import csv, sys
import numpy as np
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ie = hdr.index('Instructions Executed')
ex = np.array([float(r[ie] or 0) for r in data])
n = len(ex)
W = float(sys.argv[2]) if len(sys.argv) > 2 else np.sort(ex)[-200]  # warp steps ~ count of an always-executed instruction
print(f"static {n} instr = {n*16/1024:.1f} KB; executed {ex.sum():.3e}; warp steps ~{W:.3e}; instr/warp-step {ex.sum()/W:.0f}")
nl = (n + 7) // 8
line = np.zeros(nl)
for i in range(n): line[i // 8] = max(line[i // 8], ex[i])
for thr in (0.5, 0.25, 0.1, 0.05, 0.02, 0.01, 0.001):
    hot = ex >= thr * W
    print(f"  executed in >= {thr:5.3f} of warp steps: {hot.sum():5d} instr {hot.sum()*16/1024:5.1f} KB; lines touched {int((line >= thr*W).sum()):4d} = {(line >= thr*W).sum()*128/1024:5.1f} KB; share of executed {ex[hot].sum()/ex.sum():.3f}")
print(f"  line fetch demand per warp step if nothing were cached: {line.sum()/W:.0f} lines")
