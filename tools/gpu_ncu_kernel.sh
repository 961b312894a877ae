#!/bin/bash
# ncu --set full of one kernel (regex $KERNEL) in a reduced c2 run
mkdir -p gpurun_out
CMD="python bench.py --workload ${WL:-c2} --clips ${CLIPS:-1500} --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_ncu3.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/plain_ncu3.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"${KERNEL:-iir_overlap}" -s ${SKIP:-3} -c 1 -o gpurun_out/${OUT:-prof_iir_overlap} -f $CMD > gpurun_out/ncu_full3.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full3.log
