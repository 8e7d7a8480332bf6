#!/bin/bash
# round-2 profiling pass: ncu launch list of the default bench command + ncu --set full of the K2 cluster search,
# then the default bench line (clean, not under a profiler)
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
  --log-file gpurun_out/r2_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grid_candidates_cluster -s 4 -c 1 \
  -o gpurun_out/r2_prof_k_grid_candidates_cluster -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_k2.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -n 4 gpurun_out/r2_bench_default.err
ls -la gpurun_out/r2_prof_k_grid_candidates_cluster.ncu-rep gpurun_out/r2_launches_default.csv
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_default.json").read().strip().splitlines()[-1]); print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("stage_ms"), d["e2e"].get("stage_ms"), d.get("clocks"))
print("  d2m", d["dist2mat"]["value"], d["dist2mat"]["e2e"]["value"], d["dist2mat"]["by_face"]["e2e"]["value"]); print("  shim", d["e2e_shim"].get("call_ms")); print("  cpu", d["cpu_baseline"]["value"])
PY
