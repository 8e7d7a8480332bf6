#!/bin/bash
# round 2, first GPU pass: whole GPU test suite (new large-size parity + flagged class + reference CUDA build),
# the default bench line, the reference arm, the ncu launch list and one full capture of k_clip
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
( time timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=15 ) > gpurun_out/r2a_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2a_pytest_gpu.log | tail -5
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err; tail -n 5 gpurun_out/r2a_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r2a_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_clip -s 4 -c 1 \
  -o gpurun_out/r2a_prof_k_clip -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_k_clip.log 2>&1
cut -c1-3000 gpurun_out/r2a_bench_default.json; cut -c1-1500 gpurun_out/r2a_bench_reference.json
tail -n 40 gpurun_out/r2a_pytest_gpu.log
