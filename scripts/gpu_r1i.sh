#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|FAILED" gpurun_out/pytest_gpu.log | head -20
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_default.json")); print("default: value %.1fM e2e %.1fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), d["e2e"]["stage_ms"], d["stage_ms"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
    m=d["dist2mat"]; print("d2m", m["value"]/1e6, m["ms_per_step"], "e2e", m["e2e"]["value"]/1e6, "shared", m.get("shared_lists"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_default.err").read()[-2500:])
PY
