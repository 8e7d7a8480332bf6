#!/bin/bash
# round-2 profiling pass: ncu launch list of the default bench command + full captures of the top kernels, then the
# default bench line and the reference arm (clean, not under a profiler)
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/r2_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_clip_tiny -s 4 -c 1 \
  -o gpurun_out/r2_prof_k_clip_tiny -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_k_clip.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grid_candidates -s 4 -c 1 \
  -o gpurun_out/r2_prof_k_grid_candidates -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_k2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dist2mat_q -s 2 -c 1 \
  -o gpurun_out/r2_prof_k_dist2mat_q -f python bench.py --workload d2m --samples 2000000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_d2m.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -n 4 gpurun_out/r2_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg4_1gpu.json 2> gpurun_out/r2_bench_cfg4_1gpu.err
timeout 600 python bench.py --steps 10 --warmup 3 --mode given --no-cpu-baseline --grid-candidates > gpurun_out/r2_bench_cfg2_given.json 2> gpurun_out/r2_bench_cfg2_given.err
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err
ls -la gpurun_out/r2_*.ncu-rep
python - <<PY
import json
for f in ("r2_bench_default","r2_bench_reference","r2_bench_cfg4_1gpu","r2_bench_cfg2_given","r2_bench_cfg5"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("stage_ms"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
