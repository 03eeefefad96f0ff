#!/bin/bash
for GB in 24 12 6 3; do
  echo "=== TDC_MAX_WORKSPACE_GB=$GB"
  TDC_MAX_WORKSPACE_GB=$GB timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print(d['ms_per_step'], {k:round(v,1) for k,v in d['path']['kernel_ms_per_step'].items()}, d['gpu_launches'], d['clocks']['sm_mhz'])"
done
