#!/bin/bash
# Internal row batch vs the fused GEMM+LN kernel's cluster rounds (22 clusters x 256 query tokens = 352 rows per round):
#   gpurun --timeout 900 -- 'bash tools/gpu_batch_align.sh'
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-qformer-only --unfolded-steps 0 --parity-rows 0"
pick='import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(round(d["ms_per_step"],2),{k:round(v,2) for k,v in d["path"]["kernel_ms_per_step"].items()},d["clocks"]["sm_mhz"])'
for rep in 1 2; do
echo "== rep $rep"
echo -n "nb 1800 (default, 11 GB):  "; $B 2>/dev/null | python -c "$pick"
echo -n "nb 1760 = 5 x 352:         "; TDC_FRAMES_BATCH=1760 $B 2>/dev/null | python -c "$pick"
echo -n "nb 2112 = 6 x 352 (14 GB): "; TDC_MAX_WORKSPACE_GB=14 TDC_FRAMES_BATCH=2112 $B 2>/dev/null | python -c "$pick"
echo -n "nb 2160 (14 GB):           "; TDC_MAX_WORKSPACE_GB=14 TDC_FRAMES_BATCH=2160 $B 2>/dev/null | python -c "$pick"
echo -n "nb 2700 (17 GB):           "; TDC_MAX_WORKSPACE_GB=17 TDC_FRAMES_BATCH=2700 $B 2>/dev/null | python -c "$pick"
echo -n "nb 3600 (23 GB):           "; TDC_MAX_WORKSPACE_GB=23 TDC_FRAMES_BATCH=3600 $B 2>/dev/null | python -c "$pick"
done
