#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_rpd.py tests/test_shim.py -m gpu -q ) > gpurun_out/r2q_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2q_pytest_gpu.log | tail -3
grep -E "^E  |FAILED|ERROR" gpurun_out/r2q_pytest_gpu.log | cut -c1-300 | tail -20
timeout 900 python bench.py --no-ref-cfg2 > gpurun_out/r2q_bench_default.json 2> gpurun_out/r2q_bench_default.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2q_bench_default.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"])
print(json.dumps(d["dist2mat"].get("by_face"))[-700:])
PY
