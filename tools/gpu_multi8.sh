#!/bin/bash
mkdir -p gpurun_out
run() { # n tag env...
  n=$1; tag=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline ${EXTRA} > gpurun_out/scale_n${n}_$tag.json 2> gpurun_out/scale_n${n}_$tag.err
  echo "n=$n $tag rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/scale_n${n}_$tag.err | tail -2 | cut -c1-300
  python - <<PY
import json
for line in open('gpurun_out/scale_n${n}_$tag.json'):
    if line.startswith('{'):
        d=json.loads(line); print('n',d['n_gpus'],'$tag','value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',d['e2e'] and round(d['e2e']['value']), d['roofline']['kernels_ms_per_launch'])
PY
}
run 8 default A=1
run 8 ctas8 NCCL_MAX_CTAS=8
run 8 ctas4 NCCL_MAX_CTAS=4
