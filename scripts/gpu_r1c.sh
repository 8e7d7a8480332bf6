#!/bin/bash
# round-1c: streamed e2e check + candidate-kernel scaling probe
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for C in 0 2 8; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --chunks $C > gpurun_out/bench_c$C.json 2> gpurun_out/bench_c$C.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_c$C.json"))
    print("chunks $C", "value %.1fM"%(d["value"]/1e6), "e2e %.1fM"%(d["e2e"]["value"]/1e6), d["e2e"].get("stage_ms"), d["e2e"]["path"][:60], d["stage_ms"])
except Exception as e:
    print("chunks $C failed", e); print(open("gpurun_out/bench_c$C.err").read()[-1500:])
PY
done
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
cat gpurun_out/bench_cfg4.json; tail -3 gpurun_out/bench_cfg4.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grid_candidates -s 6 -c 1 \
  -o gpurun_out/prof_cand_cfg4 -f python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cand_cfg4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grid_candidates -s 6 -c 1 \
  -o gpurun_out/prof_cand_cfg2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cand_cfg2.log 2>&1
ls -la gpurun_out/*.ncu-rep
