#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r3d_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r3d_pytest_gpu.log | tail -3
grep -E "^E  |FAILED|ERROR" gpurun_out/r3d_pytest_gpu.log | cut -c1-300 | tail -20
grep -E "s call|s setup" gpurun_out/r3d_pytest_gpu.log | head -8
timeout 900 python bench.py > gpurun_out/r3d_bench_default.json 2> gpurun_out/r3d_bench_default.err
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r3d_bench_cfg5.json 2> gpurun_out/r3d_bench_cfg5.err
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3d_bench_cfg4.json 2> gpurun_out/r3d_bench_cfg4.err
python - <<PY
import json
for f in ("r3d_bench_default","r3d_bench_cfg5","r3d_bench_cfg4"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("stage_ms"), d["e2e"].get("stage_ms"), d.get("incremental"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
