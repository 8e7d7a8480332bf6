#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err
MB_TRACE=2 timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2r_bench_n2_trace.json 2> gpurun_out/r2r_bench_n2_trace.err
python - <<PY
import json
for f in ("r2r_bench_n2","r2r_bench_n2_trace"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("stage_ms"), d["e2e"].get("stage_ms"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
grep -c . gpurun_out/r2r_bench_n2_trace.err
