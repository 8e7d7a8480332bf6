#!/bin/bash
# N-GPU call (gpurun --gpus N): GPU tests, N=1 bench line, then the torchrun weak-scaling lines up to N
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
for G in 2 4 8; do
  [ $G -le $N ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29500+G)) \
     bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/bench_n$G.json 2> gpurun_out/bench_n$G.err
  echo "N=$G exit $?"
done
tail -4 gpurun_out/pytest_gpu.log 2>/dev/null
for G in 1 2 4 8; do [ -f gpurun_out/bench_n$G.json ] && { cat gpurun_out/bench_n$G.json; tail -n 5 gpurun_out/bench_n$G.err; }; done
