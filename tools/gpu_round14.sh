#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py -x -q 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none -k regex:tdc_gemm_kernel -s 63 -c 6 -o gpurun_out/prof_query_gemm_5400 \
   python bench.py --steps 1 --warmup 1 --segments 1800 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1; echo "ncu qgemm rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:layernorm_kernel -s 31 -c 2 -o gpurun_out/prof_layernorm_5400 \
   python bench.py --steps 1 --warmup 1 --segments 1800 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full4.log 2>&1; echo "ncu ln rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --num-text 32 --no-e2e > gpurun_out/bench14_text32.json 2> gpurun_out/bench14_text32.err; echo "bench text rc=$?"; tail -3 gpurun_out/bench14_text32.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench14_text32.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['path'], d['clocks'], d.get('cpu_baseline'))
PY
