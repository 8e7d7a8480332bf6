#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/r2p_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2p_pytest_gpu.log | tail -3
grep -E "^E  |FAILED|ERROR" gpurun_out/r2p_pytest_gpu.log | cut -c1-300 | tail -20
timeout 900 python bench.py --workload d2m --samples 10000000 --steps 10 --warmup 3 > gpurun_out/r2p_bench_d2m.json 2> gpurun_out/r2p_bench_d2m.err
timeout 900 python bench.py > gpurun_out/r2p_bench_default.json 2> gpurun_out/r2p_bench_default.err
python - <<PY
import json
for f in ("r2p_bench_d2m","r2p_bench_default"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], d.get("stage_ms"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
