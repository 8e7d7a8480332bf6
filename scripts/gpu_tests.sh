#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|FAILED" gpurun_out/pytest_gpu.log | head -30
