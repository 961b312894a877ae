#!/bin/bash
mkdir -p gpurun_out
for n in ${NS:-8 4}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  echo "n=$n rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/scale_n$n.err | tail -3 | cut -c1-300
  python - <<PY
import json
for line in open('gpurun_out/scale_n$n.json'):
    if line.startswith('{'):
        d=json.loads(line); print('n',d['n_gpus'],'value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',d['e2e'] and round(d['e2e']['value']), d['roofline']['kernels_ms_per_launch'])
PY
done
