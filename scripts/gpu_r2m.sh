#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_stream.py tests/test_gpu_shardsink.py tests/test_gpu_loop.py -m gpu -q ) > gpurun_out/r2m_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2m_pytest_gpu.log | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err
python - <<PY
import json
for f in ("r2m_bench_n1","r2m_bench_n2"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "value %.1fM (%.3f ms) e2e %.1fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6), d["e2e"]["stage_ms"], d["stage_ms"], d["e2e"]["d2h_gbs"], d["e2e"]["path"][:90])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
