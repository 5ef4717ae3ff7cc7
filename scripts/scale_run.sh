#!/bin/bash
# bench.py at N GPUs the way the driver launches it; output gpurun_out/<tag>_scale_n<N>.json
N=$1; tag=${2:-r02}
if [ "$N" = "1" ]; then python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --quick > gpurun_out/${tag}_scale_n1.json 2> gpurun_out/${tag}_scale_n1.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${tag}_scale_n$N.json 2> gpurun_out/${tag}_scale_n$N.err; fi
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_scale_n$N.json").read().strip().splitlines()[-1])
c=d.get("c4") or {}
print("N=%d value %.4g ms/step %.3f exchange %.3f ms digest_equal %s | c4 value %.4g ms/step %.2f digest %s equal %s per-round %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d.get("exchange_ms_per_round",0), d["tree_after_one_round"]["digest_equal_on_all_ranks"], c.get("value",0), c.get("ms_per_step",0), c.get("tree_digest"), c.get("digest_equal_on_all_ranks"), c.get("ms_per_round_rank0")))
PY
