#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./build/microbench > gpurun_out/microbench.json 2> gpurun_out/microbench.err; cat gpurun_out/microbench.json
CMD="python bench.py --workload c2 --clips 700 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hmfe" -c 80 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_launch_c2.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"iir_chunk" -s 6 -c 2 -o gpurun_out/prof_iir -f $CMD > gpurun_out/ncu_full_iir.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full_iir.log
