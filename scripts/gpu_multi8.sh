#!/bin/bash
# 8-GPU call: weak-scaling lines N=4, 8 (p2p gather) + N=8 NCCL gather for comparison
set -u
mkdir -p gpurun_out
for G in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29520+G)) \
     bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/bench_n$G.json 2> gpurun_out/bench_n$G.err
  echo "N=$G exit $?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 \
     bench.py --gpus 8 --steps 10 --warmup 3 --gather nccl > gpurun_out/bench_n8_nccl.json 2> gpurun_out/bench_n8_nccl.err
python - <<PY
import json
for f in ("bench_n4","bench_n8","bench_n8_nccl"):
    try:
        txt=[l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]
        d=json.loads(txt); print(f, "value %.1fM (%.3f ms) e2e %.1fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6), d["e2e"]["stage_ms"], d["stage_ms"])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-2500:])
PY
