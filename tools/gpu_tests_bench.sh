#!/bin/bash
# tests + c2/c1 bench (no CPU baseline) for a quick look at the kernel split
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-c2}; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w rc=$?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$w.json'))
    print('value',round(d['value']),'clips/s  ms/step',round(d['ms_per_step'],3),'frames/s',round(d['frames_per_s']),'e2e',d['e2e'] and (round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1)))
    print(' roofline',d['roofline']['kernel'],round(d['roofline']['frac'],3),d['roofline']['kernels_ms_per_launch'])
except Exception as e: print('ERR',e)
PY
  tail -5 gpurun_out/bench_$w.err
done
