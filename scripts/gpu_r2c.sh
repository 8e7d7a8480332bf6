#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/dev_d2m_ref.py > gpurun_out/r2c_d2m_ref.log 2>&1; tail -n 12 gpurun_out/r2c_d2m_ref.log
( timeout 900 python -m pytest tests/test_gpu_rpd.py tests/test_flagged.py tests/test_gpu_stream.py tests/test_gpu_parity_large.py tests/test_gpu_reference_build.py tests/test_gpu_loop.py tests/test_gpu_grid_edge.py -m gpu -q -s ) > gpurun_out/r2c_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2c_pytest_gpu.log | tail -5
for V in 0 2 1; do
MB_CLIP_VARIANT=$V timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_v$V.json 2> gpurun_out/r2c_bench_v$V.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c_bench_v$V.json').read().strip().splitlines()[-1])
print('variant $V value',d['value'],'ms',d['ms_per_step'],d['stage_ms'],'cells',d['run']['cells_per_step'])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_clip_tiny -s 4 -c 1 \
  -o gpurun_out/r2c_prof_k_clip_tiny -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_k_clip.log 2>&1
grep -E "FAILED|ERROR" gpurun_out/r2c_pytest_gpu.log | tail
