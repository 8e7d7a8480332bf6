#!/bin/bash
# 8-GPU call: weak-scaling lines N=1, 2, 4, 8 (peer-memory sink) + N=8 NCCL gather for comparison + reference arm at N=8
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
for G in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29520+G)) \
     bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/bench_n$G.json 2> gpurun_out/bench_n$G.err
  echo "N=$G exit $?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 \
     bench.py --gpus 8 --steps 10 --warmup 3 --gather nccl > gpurun_out/bench_n8_nccl.json 2> gpurun_out/bench_n8_nccl.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
     bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_n8_reference.json 2> gpurun_out/bench_n8_reference.err
python - <<PY
import json
for f in ("bench_n1","bench_n2","bench_n4","bench_n8","bench_n8_nccl","bench_n8_reference"):
    try:
        lines=open("gpurun_out/%s.json"%f).read().strip().splitlines()
        d=json.loads(lines[-1]); print(f, len(lines), "line(s): value %.1fM (%.3f ms) e2e %.1fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6), d["e2e"].get("stage_ms"), d.get("stage_ms"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-2500:])
PY
