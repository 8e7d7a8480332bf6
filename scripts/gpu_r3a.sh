#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_loop.py tests/test_gpu_shardsink.py tests/test_gpu_bgeo.py tests/test_shim.py -m gpu -q ) > gpurun_out/r3a_pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/r3a_pytest_gpu.log | tail -3
grep -E "^E  |FAILED|ERROR" gpurun_out/r3a_pytest_gpu.log | cut -c1-300 | tail -20
for c in 0 3 4 6 8; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --chunks $c > gpurun_out/r3a_bench_c$c.json 2> gpurun_out/r3a_bench_c$c.err
done
MB_STREAM_VARIANT=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --chunks 3 > gpurun_out/r3a_bench_c3_old.json 2> gpurun_out/r3a_bench_c3_old.err
python - <<PY
import json
for c in ("c0","c3","c4","c6","c8","c3_old"):
    f="r3a_bench_%s"%c
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); e=d["e2e"]; print(f, "value %.1f M"%(d["value"]/1e6), "e2e %.1f M"%(e["value"]/1e6), e["stage_ms"], "full %.1f M"%(e["full_records"]["value"]/1e6), e["path"][40:110])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
