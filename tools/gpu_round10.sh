#!/bin/bash
mkdir -p gpurun_out
run() { echo "--- $*"; timeout 90 ./build/gemm_test "$@"; echo "exit=$?"; }
{ run 2 1000 776 1152 2; run 1 333 136 96 0; run 2 86400 768 768 2 20; run 2 148992 9216 3584 0 5; } 2>&1 | grep -E "^---|verify|time|exit=[1-9]"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench10_n1.json 2> gpurun_out/bench10_n1.err; echo "bench n1 rc=$?"; tail -3 gpurun_out/bench10_n1.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench10_n1.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['path']['kernel_ms_per_step'], d['clocks'], d['roofline']['achieved'])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdc_attention_kernel -s 18 -c 2 -o gpurun_out/prof_attention_v4 \
   python bench.py --steps 1 --warmup 1 --segments 600 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; echo "ncu attn rc=$?"
