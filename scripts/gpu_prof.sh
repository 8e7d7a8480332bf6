#!/bin/bash
# ncu: launch list of one bench run + full capture of one kernel (regex in $1, default k_clip)
set -u
K=${1:-k_clip}
MODE=${2:-grid}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_$MODE.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --mode $MODE > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 \
  -o gpurun_out/prof_$K -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --mode $MODE > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
