#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_reference_build.py tests/test_gpu_dist2mat.py tests/test_gpu_stream.py -m gpu -q -s ) > gpurun_out/r2f_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2f_pytest_gpu.log | tail -5
grep -E "FAILED|ERROR|bit-identical|parity|lattice" gpurun_out/r2f_pytest_gpu.log | cut -c1-400 | tail -20
timeout 300 python bench.py --workload d2m --samples 2000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_d2m.json 2> gpurun_out/r2f_bench_d2m.err; cut -c1-300 gpurun_out/r2f_bench_d2m.json
