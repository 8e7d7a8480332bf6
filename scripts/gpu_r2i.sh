#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dist2mat_lists.py tests/test_gpu_dist2mat.py -m gpu -q -s ) > gpurun_out/r2i_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2i_pytest_gpu.log | tail -5
grep -E "^E  |FAILED|ERROR" gpurun_out/r2i_pytest_gpu.log | cut -c1-300 | tail -30
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 600 python bench.py --workload d2m --samples 10000000 --steps 5 --warmup 3 ) > gpurun_out/r2i_bench_d2m.json 2> gpurun_out/r2i_bench_d2m.err; tail -4 gpurun_out/r2i_bench_d2m.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2i_bench_d2m.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'])
print('by_face',d.get('by_face'))
print('gpu_reference',d.get('gpu_reference'))
PY
