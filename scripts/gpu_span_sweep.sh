#!/bin/bash
# usage: gpu_span_sweep.sh N  -- default bench + span sweep in one process, then a 3-step trace run
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
MB_TRACE=2 timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --sweep-chunks ${SWEEP:-1,2,3,4,6} > gpurun_out/sweep_bench_n$N.json 2> gpurun_out/sweep_bench_n$N.err
python - <<PY
import json
f="sweep_bench_n$N"
try:
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("stage_ms"), d["e2e"].get("stage_ms"), d.get("span_sweep_wall_ms"))
except Exception as e:
    print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
grep -c "trace" gpurun_out/sweep_bench_n$N.err
