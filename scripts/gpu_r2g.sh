#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_reference_build.py tests/test_gpu_dist2mat.py tests/test_gpu_stream.py tests/test_gpu_bgeo.py tests/test_gpu_shardsink.py tests/test_gpu_parity_large.py::test_dist2mat_10M_sampled_against_reference -m gpu -q -s ) > gpurun_out/r2g_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2g_pytest_gpu.log | tail -5
grep -E "FAILED|ERROR|bit-identical|samples:|fixture|10M" gpurun_out/r2g_pytest_gpu.log | cut -c1-400 | tail -20
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],d['stage_ms'],'e2e',d['e2e']['value'],d['e2e']['stage_ms'],d['e2e']['d2h_bytes_per_step'],'full',d['e2e']['full_records'])
PY
tail -3 gpurun_out/r2g_bench.err
