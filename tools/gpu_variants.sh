#!/bin/bash
# build and time variants of the IIR overlap kernel: "ring ctas" pairs
mkdir -p gpurun_out
for v in "$@"; do
  set -- $v
  ring=$(echo $v | cut -d: -f1); ctas=$(echo $v | cut -d: -f2)
  sed -i "s/^constexpr int kRing = [0-9]*;/constexpr int kRing = $ring;/; s/__launch_bounds__(kIirWarps \* 32, [0-9]*) iir_overlap4_kernel/__launch_bounds__(kIirWarps * 32, $ctas) iir_overlap4_kernel/; s/const double slots_o = (double)ctx->sm_count \* [0-9]* \* kIirWarps/const double slots_o = (double)ctx->sm_count * $ctas * kIirWarps/" heart_murmur_detection_b200/csrc/iir.cu
  python -m heart_murmur_detection_b200.build > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/var_$ring_$ctas.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/var_$ring_$ctas.json'))
print("ring $ring ctas $ctas:", d['config'].get('iir_plan'), d['roofline']['kernels_ms_per_launch'].get('iir_overlap'), 'step', round(d['ms_per_step'],3))
PY
done
