#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s --durations=12 ) > gpurun_out/r2b_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2b_pytest_gpu.log | tail -5
( time timeout 900 python bench.py --steps 10 --warmup 3 --no-ref-cfg2 --samples 1000000 ) > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err; tail -n 3 gpurun_out/r2b_bench_default.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_clip -s 4 -c 1 \
  -o gpurun_out/r2b_prof_k_clip -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_k_clip.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_default.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],d['stage_ms'],'e2e',d['e2e']['value'])
print('shim',d.get('e2e_shim'))
PY
grep -E "FAILED|ERROR|passed|failed" gpurun_out/r2b_pytest_gpu.log | tail -20
