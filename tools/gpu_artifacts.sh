#!/bin/bash
# Round artefacts: GPU tests, default bench line, reference arm, ncu launch list of the same command,
# ncu --set full of the step's kernels (one launch each, at the full workload size).
mkdir -p gpurun_out
T=${TAG:-r01}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$T.log
timeout 900 python bench.py > gpurun_out/BENCH_$T.json 2> gpurun_out/BENCH_$T.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 > gpurun_out/BENCH_${T}_reference.json 2> gpurun_out/BENCH_${T}_reference.err; echo "ref rc=$?"
for w in c1 c3 c2nf; do
  timeout 900 python bench.py --workload $w > gpurun_out/BENCH_${T}_$w.json 2> gpurun_out/BENCH_${T}_$w.err; echo "bench $w rc=$?"
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$T.log 2>&1 || { echo "plain failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"iir_|logmel_|trim_|gather_|fbank_|pcm16_|spec_|resample_" -c 400 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_launch_$T.log 2>&1; echo "launch list rc=$?"
# kernels of one c2 step: init_stats, iir_overlap4, trim_index_hop4, gather, logmel_power, logmel_finalize -> skip the 4 earlier steps' launches
ncu --set full --clock-control none --import-source on -k regex:"iir_overlap|logmel_power|logmel_finalize|gather_kernel|trim_index" -s 20 -c 5 -o gpurun_out/prof_${T}_c2 -f $CMD > gpurun_out/ncu_full_$T.log 2>&1; echo "full rc=$?"
CMD3="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"fbank_kernel" -s 4 -c 1 -o gpurun_out/prof_${T}_c3 -f $CMD3 > gpurun_out/ncu_full_${T}_c3.log 2>&1; echo "full c3 rc=$?"
CMD1="python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"logmel_power" -s 4 -c 1 -o gpurun_out/prof_${T}_c1 -f $CMD1 > gpurun_out/ncu_full_${T}_c1.log 2>&1; echo "full c1 rc=$?"
python - <<PY
import json
for f in ('BENCH_$T','BENCH_${T}_c1','BENCH_${T}_c3'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline']['kernels_ms_per_launch'])
print(open('gpurun_out/BENCH_${T}_reference.json').read()[:300])
PY
