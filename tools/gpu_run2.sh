#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
for v in packed scalar; do
  timeout 600 python bench.py --steps 50 --warmup 5 --variant $v --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  echo "bench $v rc=$?"; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$v.json')); print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])"; tail -3 gpurun_out/bench_$v.err
done
