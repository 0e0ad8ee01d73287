#!/bin/bash
# bench.py at N = $@ GPUs of this box, back to back (own measurement; the driver computes scaling efficiency itself).
for n in "$@"; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/scale_n$n.json') if l.startswith('{')][0]); print('N=$n value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'stdout lines', sum(1 for _ in open('gpurun_out/scale_n$n.json')))"
done
