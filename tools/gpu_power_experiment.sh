#!/bin/bash
# Does cutting L2->SM operand traffic (W-multicast clusters) pay under the 1 kW power cap of the long step?
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 600 $B 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$label', round(d['ms_per_step'],1), 'ms | kv', round(d['roofline']['achieved']), 'TF |', {k: round(v,1) for k,v in d['path']['kernel_ms_per_step'].items()}, '| clk', d['clocks']['sm_mhz'])"
}
{
run "MC=1(shipped)" TDC_GEMM_MC=1
run "MC=2" TDC_GEMM_MC=2
run "MC=1(again)" TDC_GEMM_MC=1
run "MC=2(again)" TDC_GEMM_MC=2
} 2>&1 | tee gpurun_out/power_experiment.txt
