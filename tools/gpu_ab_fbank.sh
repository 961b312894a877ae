#!/bin/bash
# A/B of the fbank mel-window placement on c3 (kernel-only bench lines) with the parity tests under each setting:
# HMFE_FBANK_MEL_GROUP (alignment that fixes the trip counts) x HMFE_FBANK_MEL_PREFER (conflict-free re-assignment)
mkdir -p gpurun_out
for cfg in ${CONFIGS:-"8:0 8:16"}; do
  g=${cfg%%:*}; pr=${cfg##*:}
  export HMFE_FBANK_MEL_GROUP=$g HMFE_FBANK_MEL_PREFER=$pr
  timeout 300 python -m pytest tests/test_stages_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x -k "fbank or vggish" > gpurun_out/ab_fbank_test_${g}_$pr.log 2>&1; echo "group $g prefer $pr pytest rc=$?"; tail -1 gpurun_out/ab_fbank_test_${g}_$pr.log
  for rep in 1 2; do
    timeout 300 python bench.py --workload c3 --no-cpu-baseline --no-e2e --steps 50 > gpurun_out/ab_fbank_${g}_$pr.json 2> gpurun_out/ab_fbank_${g}_$pr.err
    python -c "
import json; d=json.load(open('gpurun_out/ab_fbank_${g}_$pr.json')); print('group $g prefer $pr', round(d['ms_per_step'],4), 'ms/step', d['roofline']['kernels_ms_per_launch'])"
  done
done
