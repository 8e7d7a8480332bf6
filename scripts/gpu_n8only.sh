#!/bin/bash
set -u
mkdir -p gpurun_out
for G in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29530+G)) \
     bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/bench_n$G.json 2> gpurun_out/bench_n$G.err
echo "N=$G exit $?"
done
python - <<PY
import json
for f in ("bench_n4","bench_n8"):
    try:
        lines=open("gpurun_out/%s.json"%f).read().strip().splitlines()
        d=json.loads(lines[-1]); print(f, len(lines), "line(s): value %.1fM (%.3f ms) e2e %.1fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6), d["e2e"].get("stage_ms"), d.get("stage_ms"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-2500:])
PY
