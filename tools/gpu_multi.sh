#!/bin/bash
mkdir -p gpurun_out
run() { # n tag args...
  n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/scale_n${n}_$tag.json 2> gpurun_out/scale_n${n}_$tag.err
  echo "n=$n $tag rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/scale_n${n}_$tag.err | tail -4 | cut -c1-400
  python - <<PY
import json
for line in open('gpurun_out/scale_n${n}_$tag.json'):
    if line.startswith('{'):
        d=json.loads(line); print('n',d['n_gpus'],'$tag','value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',d['e2e'] and round(d['e2e']['value']), d['roofline']['kernels_ms_per_launch'])
PY
}
N=${N:-2}
run $N auto --gather auto ${EXTRA}
