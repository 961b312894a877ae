#!/bin/bash
# A/B of the fbank mel-window alignment (HMFE_FBANK_MEL_GROUP) on c3, kernel-only bench lines; parity test under each
mkdir -p gpurun_out
for g in 16 8 4; do
  HMFE_FBANK_MEL_GROUP=$g timeout 300 python -m pytest tests/test_stages_gpu.py -m gpu -q -x -k "fbank or vggish" > gpurun_out/ab_fbank_test_$g.log 2>&1; echo "group $g pytest rc=$?"; tail -1 gpurun_out/ab_fbank_test_$g.log
  for rep in 1 2; do
    HMFE_FBANK_MEL_GROUP=$g timeout 300 python bench.py --workload c3 --no-cpu-baseline --no-e2e --steps 50 > gpurun_out/ab_fbank_$g.json 2> gpurun_out/ab_fbank_$g.err
    python -c "
import json; d=json.load(open('gpurun_out/ab_fbank_$g.json')); print('group $g', round(d['ms_per_step'],4), 'ms/step', d['roofline']['kernels_ms_per_launch'])"
  done
done
