#!/bin/bash
# Light refresh after a kernel change: all GPU tests, smoke(), the c3 and default bench lines, one full ncu capture
# of the fbank kernel, and the sibling front-ends' throughput sample.
mkdir -p gpurun_out
T=${TAG:-r01b}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$T.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$T.log
timeout 600 python bench.py --workload c3 > gpurun_out/BENCH_${T}_c3.json 2> gpurun_out/BENCH_${T}_c3.err; echo "bench c3 rc=$?"
timeout 300 python tools/siblings_sample.py > gpurun_out/siblings_sample.json 2> gpurun_out/siblings_sample.err; echo "siblings rc=$?"; cat gpurun_out/siblings_sample.json
timeout 900 python bench.py > gpurun_out/BENCH_$T.json 2> gpurun_out/BENCH_$T.err; echo "bench rc=$?"
CMD3="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fbank_kernel" -s 4 -c 1 -o gpurun_out/prof_${T}_c3 -f $CMD3 > gpurun_out/ncu_full_${T}_c3.log 2>&1; echo "full c3 rc=$?"
python - <<PY
import json
for f in ('BENCH_$T','BENCH_${T}_c3'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline']['kernels_ms_per_launch'])
    except Exception as e: print(f, 'ERR', e)
PY
