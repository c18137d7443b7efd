#!/bin/bash
# strong-scaling sweep on one box: scripts/scale.sh "1 2 4" tag
for n in $1; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/$2_n$n.json 2> gpurun_out/$2_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/$2_n$n.json 2> gpurun_out/$2_n$n.err
  fi
  python -c "
import json,sys
d=json.loads(open('gpurun_out/$2_n$n.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],2), {k:round(v,3) for k,v in d['stage_ms'].items()})"
done
