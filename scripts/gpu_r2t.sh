#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_rpd.py tests/test_gpu_parity_large.py tests/test_gpu_loop.py tests/test_gpu_stream.py tests/test_flagged.py -m gpu -q -k "not dist2mat" ) > gpurun_out/r2t_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2t_pytest_gpu.log | tail -3
grep -E "^E  |FAILED|ERROR" gpurun_out/r2t_pytest_gpu.log | cut -c1-300 | tail -20
for v in 0 1; do
MB_K2_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_k2v$v.json 2> gpurun_out/r2t_bench_k2v$v.err
done
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_cfg4.json 2> gpurun_out/r2t_bench_cfg4.err
python - <<PY
import json
for f in ("r2t_bench_k2v0","r2t_bench_k2v1","r2t_bench_cfg4"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("stage_ms"), d["run"]["candidate_pairs_per_step"], d["run"]["cells_per_step"])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
