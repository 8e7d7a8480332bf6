#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_emit.py tests/test_gpu_topology.py tests/test_gpu_dist2mat.py tests/test_gpu_reference_build.py tests/test_gpu_parity_large.py::test_dist2mat_10M_sampled_against_reference -m gpu -q -s ) > gpurun_out/r2h_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2h_pytest_gpu.log | tail -5
grep -E "^E  |FAILED|ERROR|bit-identical|samples:|fixture|10M" gpurun_out/r2h_pytest_gpu.log | cut -c1-300 | tail -30
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
