#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_shardsink.py -m gpu -q ) 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
  --log-file gpurun_out/r2_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ls -la gpurun_out/r2_launches_default.csv
