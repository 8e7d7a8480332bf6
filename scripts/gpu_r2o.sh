#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/r2o_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2o_pytest_gpu.log | tail -3
grep -E "^E  |FAILED|ERROR" gpurun_out/r2o_pytest_gpu.log | cut -c1-300 | tail -20
timeout 600 python bench.py --steps 10 --warmup 3 --mode given --no-cpu-baseline --grid-candidates > gpurun_out/r2o_bench_cfg2_given.json 2> gpurun_out/r2o_bench_cfg2_given.err
MB_CLIP_VARIANT=2 timeout 600 python bench.py --steps 10 --warmup 3 --mode given --no-cpu-baseline --grid-candidates > gpurun_out/r2o_bench_cfg2_given_v2.json 2> gpurun_out/r2o_bench_cfg2_given_v2.err
python - <<PY
import json
for f in ("r2o_bench_cfg2_given","r2o_bench_cfg2_given_v2"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("stage_ms"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
