#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|FAILED" gpurun_out/pytest_gpu.log | head -20
for C in 0 2 8; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --chunks $C > gpurun_out/bench_c$C.json 2> gpurun_out/bench_c$C.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_c$C.json"))
    print("chunks $C", "value %.1fM"%(d["value"]/1e6), "e2e %.1fM"%(d["e2e"]["value"]/1e6), d["e2e"].get("stage_ms"), d["stage_ms"])
except Exception as e:
    print("chunks $C failed", e); print(open("gpurun_out/bench_c$C.err").read()[-1500:])
PY
done
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_cfg4.json")); print("cfg4 value %.1fM e2e %.1fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), d["e2e"]["stage_ms"], d["stage_ms"])
PY
for V in 0 1; do
MB_D2M_VARIANT=$V timeout 600 python bench.py --workload d2m --samples 2000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d2m_v$V.json 2> gpurun_out/bench_d2m_v$V.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_d2m_v$V.json")); print("d2m variant $V: %.1fM q/s, %.3f ms, frac %.4f"%(d["value"]/1e6,d["ms_per_step"],d["roofline"]["frac"]))
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dist2mat_q -s 2 -c 1 \
  -o gpurun_out/prof_k_dist2mat_q -f python bench.py --workload d2m --samples 2000000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_d2m.log 2>&1
ls -la gpurun_out/*.ncu-rep
