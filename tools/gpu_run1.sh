#!/bin/bash
# first GPU trip: microbench, parity tests, smoke, bench (both variants)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 ./build/microbench > gpurun_out/microbench.json 2> gpurun_out/microbench.err
echo "microbench rc=$?"; cat gpurun_out/microbench.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
for v in packed scalar; do
  timeout 600 python bench.py --steps 20 --warmup 5 --variant $v > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  echo "bench $v rc=$?"; cat gpurun_out/bench_$v.json; tail -3 gpurun_out/bench_$v.err
done
