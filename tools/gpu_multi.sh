#!/bin/bash
# N-GPU bench lines (weak scaling: one c2 batch per rank + the all-gather)
mkdir -p gpurun_out
N=${N:-2}
for n in $(seq 1 $N); do
  case $n in 1|2|4|8) ;; *) continue;; esac
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline ${EXTRA} > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline ${EXTRA} > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "n=$n rc=$?"; tail -3 gpurun_out/scale_n$n.err | cut -c1-300
  python - <<PY
import json
for line in open('gpurun_out/scale_n$n.json'):
    if line.startswith('{'):
        d=json.loads(line); print('n',d['n_gpus'],'value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',d['e2e'] and round(d['e2e']['value']), d['roofline']['kernels_ms_per_launch'])
PY
done
