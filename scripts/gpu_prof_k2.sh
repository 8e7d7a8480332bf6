#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grid_candidates -s 4 -c 1 \
  -o gpurun_out/r1c_prof_k_grid_candidates -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_k2.log 2>&1
tail -3 gpurun_out/ncu_full_k2.log
ls -la gpurun_out/r1c_prof_k_grid_candidates.ncu-rep
