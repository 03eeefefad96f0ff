#!/bin/bash
mkdir -p gpurun_out
run() { echo "--- $*"; timeout 90 ./build/gemm_test "$@"; echo "exit=$?"; }
{ run 2 1000 768 1152 1; run 2 86400 3072 768 1 20; run 1 86400 3072 768 1 20; run 1 333 136 96 1; } 2>&1 | grep -E "^---|verify|time|exit=[1-9]"
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench19_n1.json 2> gpurun_out/bench19_n1.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench19_n1.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['path']['kernel_ms_per_step'], d['clocks'])
PY
