#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "^E  |passed|failed|FAILED" gpurun_out/pytest_gpu.log | head -10
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c0.json 2> gpurun_out/bench_c0.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode given --grid-candidates > gpurun_out/bench_given.json 2> gpurun_out/bench_given.err
python - <<PY
import json
for f in ("bench_c0","bench_given"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, "value %.1fM e2e %.1fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), d["e2e"]["stage_ms"], d["stage_ms"])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-2500:])
PY
