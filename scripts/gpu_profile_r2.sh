#!/bin/bash
# refreshed profiling pass after the K2 cluster search: launch list of the default bench command + full capture of K2
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/r2_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grid_candidates_cluster -s 4 -c 1 \
  -o gpurun_out/r2_prof_k_grid_candidates_cluster -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_k2.log 2>&1
ls -la gpurun_out/r2_*.ncu-rep gpurun_out/r2_launches_default.csv
