#!/bin/bash
# 2-GPU call: sink tests + N=1 / N=2 bench in both gather modes
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | head -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
for M in p2p nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus 2 --steps 10 --warmup 3 --gather $M > gpurun_out/bench_n2_$M.json 2> gpurun_out/bench_n2_$M.err
  echo "N=2 $M exit $?"
done
python - <<PY
import json
for f in ("bench_n1","bench_n2_p2p","bench_n2_nccl"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, "value %.1fM (%.3f ms) e2e %.1fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6), d["e2e"]["stage_ms"], d["stage_ms"])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-2500:])
PY
