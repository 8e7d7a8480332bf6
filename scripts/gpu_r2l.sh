#!/bin/bash
# 8-GPU call: N=4 and N=8 (= BASELINE configs[3], the north-star target) lines + the reference arm at N=8
set -u
mkdir -p gpurun_out
for G in 8 4; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29520+G)) \
     bench.py --gpus $G --steps 10 --warmup 3 ) > gpurun_out/r2l_bench_n$G.json 2> gpurun_out/r2l_bench_n$G.err
  echo "N=$G exit $?"; tail -n 4 gpurun_out/r2l_bench_n$G.err | grep real
done
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
     bench.py --impl reference --gpus 8 --steps 3 --warmup 1 ) > gpurun_out/r2l_bench_n8_reference.json 2> gpurun_out/r2l_bench_n8_reference.err
tail -n 4 gpurun_out/r2l_bench_n8_reference.err | grep real
python - <<PY
import json
for f in ("r2l_bench_n4","r2l_bench_n8","r2l_bench_n8_reference"):
    try:
        lines=open("gpurun_out/%s.json"%f).read().strip().splitlines()
        d=json.loads(lines[-1]); print(f, len(lines), "line(s): value %.1fM (%.3f ms) e2e %.1fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6), d["e2e"].get("stage_ms"), d.get("stage_ms"), d["config"]["workload"][:60], d.get("cpu_baseline",{}).get("cores"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-2500:])
PY
