"""profiles/r02_rollout_ops.json: the floating-point work the round kernel EXECUTES on the C3 workload, from ncu.

    ncu --metrics <METRICS> --clock-control none -k regex:rollout_kernel -s 2 -c 1 --csv --log-file gpurun_out/ops_fp64.csv \
        python scripts/quick_round.py            (and CLRRT_FP32=1 ... ops_fp32.csv)
    python scripts/rollout_ops.py gpurun_out/ops_fp64.csv gpurun_out/ops_fp32.csv

bench.py divides these per-launch counts by the kernel duration it measures live (roofline.achieved).  The counts belong to
the workload (same snapshot, same samples), not to the clock; speculation makes them vary by well under 1 % run to run."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,"
           "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,"
           "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,"
           "smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,"
           "sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,"
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "launch__registers_per_thread,launch__grid_size,launch__block_size")


def parse(path):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "ID")
    out = {}
    for r in rows:
        if len(r) == len(hdr) and r[0] != "ID":
            d = dict(zip(hdr, r))
            out[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
            out["kernel"] = d["Kernel Name"]
    return out


def digest(m, source):
    g = lambda k: m.get(k, 0.0)  # noqa: E731
    return {"dadd": g("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"), "dmul": g("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"),
            "dfma": g("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"), "fadd": g("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"),
            "fmul": g("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"), "ffma": g("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"),
            "inst_executed": g("smsp__inst_executed.sum"), "threads_per_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "pipe_fp64_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "pipe_fma_pct": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "gpu_time_ms": g("gpu__time_duration.sum") * 1e-6, "dram_bytes": g("dram__bytes_read.sum") + g("dram__bytes_write.sum"),
            "registers_per_thread": g("launch__registers_per_thread"), "grid": g("launch__grid_size"), "block": g("launch__block_size"),
            "kernel": m.get("kernel"), "source": source}


if __name__ == "__main__":
    if len(sys.argv) == 2 and sys.argv[1] == "--metrics":
        print(METRICS)
        sys.exit(0)
    out = {"fp64": digest(parse(sys.argv[1]), os.path.basename(sys.argv[1]))}
    if len(sys.argv) > 2:
        out["fp32"] = digest(parse(sys.argv[2]), os.path.basename(sys.argv[2]))
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_rollout_ops.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
