#!/bin/bash
# round-1c profiling: ncu launch list of the default bench command + full captures of the three top kernels
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; grep -E "passed|failed" gpurun_out/pytest_gpu.log

timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r1c_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
for K in k_clip; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 \
  -o gpurun_out/r1c_prof_$K -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$K.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grid_candidates -s 4 -c 1 -o gpurun_out/r1c_prof_k_grid_candidates -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_k2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dist2mat_q -s 2 -c 1 \
  -o gpurun_out/r1c_prof_k_dist2mat_q -f python bench.py --workload d2m --samples 2000000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_d2m.log 2>&1
timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r1c_bench_cfg5.json 2> gpurun_out/bench_cfg5.err
timeout 900 python bench.py --workload d2m --samples 10000000 --steps 5 --warmup 3 > gpurun_out/r1c_bench_d2m_10M.json 2> gpurun_out/bench_d2m.err
timeout 600 python bench.py --steps 10 --warmup 3 --mode given --no-cpu-baseline --grid-candidates > gpurun_out/r1c_bench_cfg2_given.json 2> gpurun_out/bench_given.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1c_bench_cfg2_reference.json 2> gpurun_out/bench_reference.err
ls -la gpurun_out/*.ncu-rep; tail -n 2 gpurun_out/bench_cfg4.err gpurun_out/bench_cfg5.err gpurun_out/bench_d2m.err gpurun_out/bench_given.err
cut -c1-400 gpurun_out/r1c_bench_cfg4.json gpurun_out/r1c_bench_cfg5.json gpurun_out/r1c_bench_d2m_10M.json gpurun_out/r1c_bench_cfg2_given.json gpurun_out/r1c_bench_cfg2_reference.json
