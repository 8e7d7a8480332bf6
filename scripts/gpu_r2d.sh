#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/dev_d2m_ref.py > gpurun_out/r2d_d2m_ref.log 2>&1; tail -n 12 gpurun_out/r2d_d2m_ref.log | cut -c1-600
( timeout 1200 python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/r2d_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2d_pytest_gpu.log | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 4 > gpurun_out/r2d_bench_l4.json 2> gpurun_out/r2d_bench_l4.err
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_cfg4.json 2> gpurun_out/r2d_bench_cfg4.err
python - <<PY
import json
for f in ('l4','cfg4'):
    try:
        d=json.loads(open('gpurun_out/r2d_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f,'value',d['value'],'ms',d['ms_per_step'],d['stage_ms'],'cells',d['run']['cells_per_step'],'e2e',d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
PY
tail -n 3 gpurun_out/r2d_bench_cfg4.err
grep -E "FAILED|ERROR" gpurun_out/r2d_pytest_gpu.log | tail; grep -A 16 "slowest" gpurun_out/r2d_pytest_gpu.log
