#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/dev_d2h_probe2.py 2>&1 | sed -n 1,1p
for c in 0 2 3 4 6 8 12; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --chunks $c > gpurun_out/r2v_bench_c$c.json 2> gpurun_out/r2v_bench_c$c.err
done
python - <<PY
import json
for c in (0,2,3,4,6,8,12):
    f="r2v_bench_c%d"%c
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); e=d["e2e"]; print(f, "value %.1f M"%(d["value"]/1e6), "e2e %.1f M"%(e["value"]/1e6), e["stage_ms"], "full %.1f M"%(e["full_records"]["value"]/1e6), e["path"][40:110])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
