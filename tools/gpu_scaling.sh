#!/bin/bash
# The driver's scaling run: N = 1, 2, 4, 8 back to back on one box (gpurun --gpus 8).
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N == 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "N=$N rc=$?"; python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/scale_n{n}.json') if l.startswith('{')][-1])
    print(n, round(d['value']), round(d['ms_per_step'],1), d['config'].get('exchange'), d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), d['path']['kernel_ms_per_step'])
except Exception as e:
    print("parse failed", e); print(open(f'gpurun_out/scale_n{n}.err').read()[-1500:])
PY
done
