#!/bin/bash
# Refresh of the c1 / c2 evidence for the final binaries: c1 bench line, full ncu captures of the c1 and c2 kernels,
# ncu launch list of the default command.  Each step has its own timeout; the steps are ordered by importance.
mkdir -p gpurun_out
T=${TAG:-r01}
timeout 120 python bench.py --workload c1 > gpurun_out/BENCH_${T}_c1.json 2> gpurun_out/BENCH_${T}_c1.err; echo "bench c1 rc=$?"
CMD1="python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 90 ncu --set full --clock-control none --import-source on -k regex:"logmel_power" -s 4 -c 1 -o gpurun_out/prof_${T}_c1 -f $CMD1 > gpurun_out/ncu_full_${T}_c1.log 2>&1; echo "full c1 rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"iir_overlap|logmel_power|logmel_finalize|gather_kernel|trim_index" -s 20 -c 5 -o gpurun_out/prof_${T}_c2 -f $CMD > gpurun_out/ncu_full_$T.log 2>&1; echo "full c2 rc=$?"
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"iir_|logmel_|trim_|gather_|fbank_|pcm16_|spec_|resample_" -c 400 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_launch_$T.log 2>&1; echo "launch list rc=$?"
