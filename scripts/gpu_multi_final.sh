#!/bin/bash
# final multi-GPU lines: bench.py at N ranks (clean, no trace) + the reference arm at N=8
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err
if [ "$N" = "8" ]; then
timeout 900 $TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_n${N}_reference.json 2> gpurun_out/final_bench_n${N}_reference.err
fi
python - <<PY
import json
for f in ("final_bench_n$N","final_bench_n${N}_reference"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("stage_ms"), d["e2e"].get("stage_ms"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "FAILED", e)
PY
