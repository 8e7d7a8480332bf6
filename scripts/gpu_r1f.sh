#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|FAILED" gpurun_out/pytest_gpu.log | head -20
timeout 600 python scripts/dev_trace.py 32 10000 1,2,4,8 3 2>&1 | tee gpurun_out/trace_cfg2.log
for C in 1 8; do
MB_TRACE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_stream_c$C.csv \
   python scripts/dev_trace.py 32 10000 $C 1 > gpurun_out/ncu_stream_c$C.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_stream_c$C.csv")) if len(r)>10 and r[0].isdigit()]
# last call = last launches; aggregate the final 1/5 (5 calls: run + 3 warm + 1)
per=collections.OrderedDict(); 
n=len(rows); tail=rows[-(n//5):]
tot=0
for r in tail:
    k=r[4][:40]; v=float(r[-1]); per[k]=per.get(k,0)+v; tot+=v
print("chunks $C: kernels in last call", len(tail), "sum us", tot/1e3 if tot>1e5 else tot)
for k,v in sorted(per.items(), key=lambda x:-x[1])[:8]: print("   ", k, v)
PY
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c0.json 2> gpurun_out/bench_c0.err
python - <<PY
import json
for f in ("bench_c0",):
    d=json.load(open("gpurun_out/%s.json"%f)); print(f, "value %.1fM e2e %.1fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), d["e2e"]["stage_ms"], d["stage_ms"])
PY
