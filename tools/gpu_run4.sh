#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for w in c2 c1; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w rc=$?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$w.json'))
    print('value',round(d['value']),'clips/s  ms/step',round(d['ms_per_step'],3),'frames/s',round(d['frames_per_s']),'e2e',d['e2e'] and round(d['e2e']['value']))
    print(' roofline',d['roofline']['kernel'],round(d['roofline']['frac'],3),d['roofline']['kernels_ms_per_launch'])
except Exception as e: print('ERR',e)
PY
  tail -5 gpurun_out/bench_$w.err
done
CMD="python bench.py --workload c1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_c1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:logmel_power -s 3 -c 1 -o gpurun_out/prof_logmel_p12 -f $CMD > gpurun_out/ncu_full_c1.log 2>&1
echo "ncu rc=$?"
