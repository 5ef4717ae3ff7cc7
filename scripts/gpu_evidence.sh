#!/bin/bash
# One GPU call that produces the round's evidence under gpurun_out/<tag>_*: GPU tests, both bench arms, the ncu launch list,
# the executed-FLOP counts of the round kernel (fp64 and fp32), one ncu --set full capture of rollout + nearest kernels.
#   gpurun --timeout 1800 -- 'bash scripts/gpu_evidence.sh r02a [notests]'
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
if [ "$2" != "notests" ]; then
  python -m pytest tests -q -m gpu > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
fi
# executed-FLOP counts of the round kernel first: bench.py's roofline divides them by the kernel time it measures live
M=$(python scripts/rollout_ops.py --metrics)
timeout 600 ncu --metrics $M --clock-control none -k regex:rollout_kernel -s 2 -c 1 --csv --log-file $out/${tag}_ops_fp64.csv python scripts/quick_round.py > $out/${tag}_ops_fp64.log 2>&1
CLRRT_FP32=1 timeout 600 ncu --metrics $M --clock-control none -k regex:rollout_kernel -s 2 -c 1 --csv --log-file $out/${tag}_ops_fp32.csv python scripts/quick_round.py > $out/${tag}_ops_fp32.log 2>&1
python scripts/rollout_ops.py $out/${tag}_ops_fp64.csv $out/${tag}_ops_fp32.csv > $out/${tag}_ops.log 2>&1 && cp profiles/r02_rollout_ops.json $out/${tag}_rollout_ops.json
python bench.py --impl reference --steps 10 --warmup 2 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || tail -5 $out/${tag}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --quick --no-c4 --sustain-seconds 0 > $out/${tag}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:rollout_kernel|nearest_sorted_kernel' -s 6 -c 2 -o $out/${tag}_prof -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --quick --no-c4 --sustain-seconds 0 > $out/${tag}_ncu_full.log 2>&1
ncu -i $out/${tag}_prof.ncu-rep --page raw --csv > $out/${tag}_prof_raw.csv 2>/dev/null
python -c "
import json
d=json.load(open('$out/${tag}_bench.json')); r=json.load(open('$out/${tag}_bench_reference.json'))
print('ours value %.4g e2e %.4g ms/step %.3f kernel_ms %s | reference %.4g (%d cores) | e2e ratio %.0f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline'].get('kernel_ms_per_launch'), r['value'], r['cpu_baseline']['cores'], d['e2e']['value']/r['value']))
print('k1', d.get('query_200ms_k1'), 'ref', r.get('query_200ms_k1'))
print('c5', d.get('c5'), 'ref', r.get('c5'))
"
