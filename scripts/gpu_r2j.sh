#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_loop.py -m gpu -q -s ) > gpurun_out/r2j_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2j_pytest_gpu.log | tail -5
grep -E "^E  |FAILED|ERROR|incremental" gpurun_out/r2j_pytest_gpu.log | cut -c1-300 | tail -30
( time timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 ) > gpurun_out/r2j_bench_cfg5.json 2> gpurun_out/r2j_bench_cfg5.err; tail -4 gpurun_out/r2j_bench_cfg5.err
cut -c1-2500 gpurun_out/r2j_bench_cfg5.json
