#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 0 2 3 4 5; do
MB_K2_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_k2v$v.json 2> gpurun_out/r2u_bench_k2v$v.err
done
for v in 0 3; do
MB_K2_VARIANT=$v timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_cfg4_k2v$v.json 2> gpurun_out/r2u_bench_cfg4_k2v$v.err
done
python - <<PY
import json
for f in ("r2u_bench_k2v0","r2u_bench_k2v2","r2u_bench_k2v3","r2u_bench_k2v4","r2u_bench_k2v5","r2u_bench_cfg4_k2v0","r2u_bench_cfg4_k2v3"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "value %.1f M"%(d["value"]/1e6), "ms %.3f"%d["ms_per_step"], d.get("stage_ms"), d["run"]["candidate_pairs_per_step"], d["run"]["cells_per_step"])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
