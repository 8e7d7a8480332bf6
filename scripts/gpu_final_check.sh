#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/final_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/final_pytest_gpu.log | tail -3
grep -E "^E  |FAILED|ERROR" gpurun_out/final_pytest_gpu.log | cut -c1-300 | tail -20
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py ) > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; tail -n 4 gpurun_out/final_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; tail -n 4 gpurun_out/final_bench_reference.err
python - <<PY
import json
for f in ("final_bench_default","final_bench_reference"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("stage_ms"), d["e2e"].get("stage_ms"), d.get("clocks"))
        if "dist2mat" in d: print("  d2m", d["dist2mat"]["value"], d["dist2mat"]["e2e"], d["dist2mat"]["by_face"]["e2e"])
        if "e2e_shim" in d: print("  shim", d["e2e_shim"].get("call_ms"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
