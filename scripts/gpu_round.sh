#!/bin/bash
# cfg4 / cfg5 / dist2mat-10M bench lines + ncu launch list + full captures of the three top kernels
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
timeout 900 python bench.py --workload d2m --samples 10000000 --steps 5 --warmup 3 > gpurun_out/bench_d2m.json 2> gpurun_out/bench_d2m.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_grid.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
for K in k_clip k_grid_candidates; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 \
  -o gpurun_out/prof_$K -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$K.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dist2mat -s 2 -c 1 \
  -o gpurun_out/prof_k_dist2mat -f python bench.py --workload d2m --samples 2000000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_d2m.log 2>&1
cat gpurun_out/bench_cfg4.json gpurun_out/bench_cfg5.json gpurun_out/bench_d2m.json
tail -n 3 gpurun_out/bench_cfg4.err gpurun_out/bench_cfg5.err gpurun_out/bench_d2m.err
