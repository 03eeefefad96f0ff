#!/bin/bash
# A/B of the engine's workspace cap (internal row-batch size) inside the power-capped step, alternating runs.
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() {
  local label=$1; shift
  env "$@" timeout 600 $B 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$label', round(d['ms_per_step'],1), 'ms | kv', round(d['roofline']['achieved']), 'TF |', {k: round(v,1) for k,v in d['path']['kernel_ms_per_step'].items()}, '| clk', d['clocks']['sm_mhz'], '| launches', d['gpu_launches'])"
}
{
for rep in 1 2; do
run "ws=48GB" TDC_MAX_WORKSPACE_GB=48
run "ws=24GB(shipped)" TDC_MAX_WORKSPACE_GB=24
run "ws=12GB" TDC_MAX_WORKSPACE_GB=12
run "ws=8GB" TDC_MAX_WORKSPACE_GB=8
done
} 2>&1 | tee gpurun_out/ws_ab.txt
