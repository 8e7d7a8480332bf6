#!/bin/bash
# 2-GPU call: N=1 and N=2 bench lines (peer-memory sink, slim records), N=2 with full records and NCCL gather for comparison
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err
for V in "slim p2p" "full p2p" "slim nccl"; do
  set -- $V
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus 2 --steps 10 --warmup 3 --records $1 --gather $2 > gpurun_out/r2k_bench_n2_$1_$2.json 2> gpurun_out/r2k_bench_n2_$1_$2.err
  echo "N=2 $V exit $?"
done
MB_TRACE=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
     bench.py --gpus 2 --steps 3 --warmup 3 > /dev/null 2> gpurun_out/r2k_trace_n2.err
python - <<PY
import json
for f in ("r2k_bench_n1","r2k_bench_n2_slim_p2p","r2k_bench_n2_full_p2p","r2k_bench_n2_slim_nccl"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "value %.1fM (%.3f ms) e2e %.1fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6), d["e2e"]["stage_ms"], d["stage_ms"], d["e2e"]["d2h_bytes_per_step"])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
grep "trace" gpurun_out/r2k_trace_n2.err | tail -6
