#!/bin/bash
# ncu captures: launch list + full capture of the dominant kernel (one variant per call via $1)
V=${1:-packed}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --variant $V"
$CMD > gpurun_out/plain_$V.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$V.csv $CMD > gpurun_out/ncu_launch_$V.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_$V.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:logmel_power -s 3 -c 2 -o gpurun_out/prof_logmel_$V -f $CMD > gpurun_out/ncu_full_$V.log 2>&1
echo "full rc=$?"; tail -5 gpurun_out/ncu_full_$V.log
