#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -s --durations=8 ) > gpurun_out/r2e_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2e_pytest_gpu.log | tail -5
grep -E "FAILED|ERROR|bit-identical|parity|lattice" gpurun_out/r2e_pytest_gpu.log | cut -c1-400 | tail -20
