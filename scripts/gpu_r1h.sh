#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|FAILED" gpurun_out/pytest_gpu.log | head -20
MB_TRACE=0 timeout 600 python scripts/dev_trace.py 32 10000 1,4,6,8 5 2>&1 | tee gpurun_out/trace_cfg2.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c0.json 2> gpurun_out/bench_c0.err
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
python - <<PY
import json
for f in ("bench_c0","bench_cfg4"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, "value %.1fM e2e %.1fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), d["e2e"]["stage_ms"], d["stage_ms"])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-2500:])
PY
