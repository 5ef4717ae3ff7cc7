#!/bin/bash
# GPU experiment: round-kernel variants (see scripts/build_variant.sh); output in gpurun_out/exp1_*.log
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/exp1_gpu.log 2>&1
for v in base t256 t256b1 t256b2 b1 mb3 sdiv t512b2; do
  CLRRT_LIB=$PWD/variants/$v.so timeout 300 python scripts/quick_round.py >> gpurun_out/exp1_quick.log 2>&1 || echo "$v FAILED" >> gpurun_out/exp1_quick.log
done
CLRRT_LIB=$PWD/variants/base.so timeout 300 python scripts/sweep_tuning.py > gpurun_out/exp1_sweep.log 2>&1
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warp_latency_per_inst_issued.ratio
for v in base t256b1 t256b2; do
  CLRRT_LIB=$PWD/variants/$v.so timeout 400 ncu --metrics $M --clock-control none -k regex:rollout_kernel -s 2 -c 1 --csv --log-file gpurun_out/exp1_ncu_$v.csv python scripts/quick_round.py > gpurun_out/exp1_ncu_$v.log 2>&1
done
cat gpurun_out/exp1_quick.log
