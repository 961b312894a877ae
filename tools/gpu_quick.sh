#!/bin/bash
# quick check after a kernel change: all GPU tests, sibling throughput sample, kernel-only bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_quick.log
timeout 300 python tools/siblings_sample.py > gpurun_out/siblings_sample.json 2> gpurun_out/siblings_sample.err; echo "siblings rc=$?"; cat gpurun_out/siblings_sample.json
for w in ${WORKLOADS:-c3}; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/quick_$w.json 2> gpurun_out/quick_$w.err; echo "bench $w rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/quick_$w.json')); print('$w', round(d['value']), 'clips/s', round(d['ms_per_step'],4), 'ms/step', d['roofline']['kernels_ms_per_launch'])"
done
