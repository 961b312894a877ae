#!/bin/bash
# Round artefacts: default bench line, reference arm, ncu launch list of the same command, ncu --set full of the top kernels.
mkdir -p gpurun_out
T=${TAG:-r01}
timeout 900 python bench.py > gpurun_out/BENCH_$T.json 2> gpurun_out/BENCH_$T.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 > gpurun_out/BENCH_${T}_reference.json 2> gpurun_out/BENCH_${T}_reference.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$T.log 2>&1 || { echo "plain failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_launch_$T.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"iir_overlap|logmel_power|logmel_finalize|gather_kernel" -s 12 -c 4 -o gpurun_out/prof_${T}_c2 -f $CMD > gpurun_out/ncu_full_$T.log 2>&1; echo "full rc=$?"
tail -2 gpurun_out/ncu_full_$T.log
python - <<PY
import json
d=json.load(open('gpurun_out/BENCH_$T.json'))
print({k:d[k] for k in ('value','ms_per_step','frames_per_s')}, d['e2e'], d['cpu_baseline'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['kernels_ms_per_launch'])
print(open('gpurun_out/BENCH_${T}_reference.json').read()[:600])
PY
