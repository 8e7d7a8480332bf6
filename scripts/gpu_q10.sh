#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_occ.json 2> gpurun_out/bench_occ.err
timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_occ4.json 2> gpurun_out/bench_occ4.err
python - <<PY
import json
for f in ("bench_occ","bench_occ4"):
    d=json.load(open("gpurun_out/%s.json"%f)); print(f, "value %.1fM"%(d["value"]/1e6), d["stage_ms"], "e2e %.1fM"%(d["e2e"]["value"]/1e6))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_clip -s 4 -c 1 \
  -o gpurun_out/r1c_prof_k_clip -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_k_clip.log 2>&1
ls -la gpurun_out/r1c_prof_k_clip.ncu-rep
