#!/bin/bash
# tests + smoke + the three bench lines (no profiler)
set -u
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_grid.json 2> gpurun_out/bench_grid.err
timeout 600 python bench.py --steps 10 --warmup 3 --mode given --no-cpu-baseline --grid-candidates > gpurun_out/bench_given.json 2> gpurun_out/bench_given.err
tail -15 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_grid.json gpurun_out/bench_given.json; tail -n 3 gpurun_out/bench_grid.err gpurun_out/bench_given.err
