#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|FAILED" gpurun_out/pytest_gpu.log | head -20
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r1c_bench_cfg5.json 2> gpurun_out/bench_cfg5.err
cut -c1-700 gpurun_out/r1c_bench_cfg5.json; tail -3 gpurun_out/bench_cfg5.err
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
