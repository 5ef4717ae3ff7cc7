"""Turn gpurun_out/launches.csv and gpurun_out/prof.ncu-rep into the tracked summaries under profiles/."""
import collections, csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
pre = (sys.argv[2] + "_") if len(sys.argv) > 2 else ""   # gpurun_out/<pre>launches.csv, gpurun_out/<pre>prof.ncu-rep
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", pre + "launches.csv"))) if len(r) > 10]
hdr = rows[0]; ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ik].split("(")[0][:70], []).append(float(r[iv].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(out, f"{tag}_launches.md"), "w") as f:
    f.write(f"# ncu launch list ({tag}): `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --quick --no-c4 --sustain-seconds 0`\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch device time (cold-cache, serialised: compare shares).\n"
            "Includes the warm-up rounds, the e2e leg (export_nodes) and the one extra round whose tree is digested.\n\n| kernel | launches | total ms | max ms | share |\n|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{k}` | {len(v)} | {sum(v)/1e6:.3f} | {max(v)/1e6:.3f} | {100*sum(v)/tot:.1f}% |\n")
print(open(os.path.join(out, f"{tag}_launches.md")).read())
rep = os.path.join(ROOT, "gpurun_out", pre + "prof.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h = r[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_xu.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        # instruction fetch: L1.5 (ICC, per SM) hit rate and requests, requests that reach the GPC-level cache (GCC)
        "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum", "gcc__cache_requests_type_instruction.sum",
        "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed"]
with open(os.path.join(out, f"{tag}_rollout_kernel_ncu.md"), "w") as f:
    f.write(f"# ncu --set full ({tag}): candidate search and rollout kernel of one C3 round (65536 samples, 1000 obstacles, 4096-node tree)\n\n"
            "`ncu --set full --clock-control none --import-source on -k 'regex:rollout_kernel|nearest_sorted_kernel' -s 6 -c 2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --quick --no-c4 --sustain-seconds 0`\n\n")
    cols = [i for i in range(2, len(r))]
    f.write("| metric | unit | " + " | ".join(f"launch {i-2}" for i in cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
    for w in want:
        if w in h:
            i = h.index(w)
            f.write(f"| {w} | {r[1][i]} | " + " | ".join(r[c][i][:60] for c in cols) + " |\n")
print(open(os.path.join(out, f"{tag}_rollout_kernel_ncu.md")).read())

# digest for bench.py (roofline.traffic and the executed-instruction view of the rollout kernel)
import json
def col_of(name_part):
    ik = h.index("Kernel Name")
    for c in range(2, len(r)):
        if name_part in r[c][ik]:
            return c
    return None
c = col_of("rollout_kernel")
if c is not None:
    def val(m):
        return float(r[c][h.index(m)].replace(",", "")) if m in h and r[c][h.index(m)] not in ("", "n/a") else None
    def unit(m):
        return r[1][h.index(m)] if m in h else ""
    def to_bytes(m):
        v, u = val(m), unit(m).lower()
        if v is None: return None
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    digest = {
        "source": f"profiles/{tag}_rollout_kernel_ncu.md (ncu --set full, one C3 round)",
        "kernel": r[c][h.index("Kernel Name")][:60],
        "duration_ms_under_ncu": val("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1, "ms": 1, "nsecond": 1e-6}.get(unit("gpu__time_duration.sum"), 1),
        "dram_bytes_per_launch": (to_bytes("dram__bytes_read.sum") or 0) + (to_bytes("dram__bytes_write.sum") or 0),
        "issue_slots_busy_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "threads_per_warp_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "warp_instructions": val("smsp__inst_executed.sum"),
        "icc_hit_rate_pct": val("sm__icc_request_hit_rate.pct"),
        "registers_per_thread": val("launch__registers_per_thread"),
    }
    json.dump(digest, open(os.path.join(out, "rollout_kernel_ncu_latest.json"), "w"), indent=1)
    print(json.dumps(digest, indent=1))
