#!/bin/bash
mkdir -p gpurun_out
python tools/audio_dbg.py 2>&1 | tail -5
run() { echo "--- $*"; timeout 90 ./build/gemm_test "$@"; echo "exit=$?"; }
{ run 2 1000 768 1152 1; run 2 86400 3072 768 1 20; run 2 86400 768 768 2 20; run 2 86400 768 768 0 20; run 2 86400 2304 768 0 20; run 2 86400 768 3072 2 20; run 2 148992 9216 3584 0 5; run 1 1000 768 1152 2; } 2>&1 | grep -E "^---|verify|time|exit=[1-9]"
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench16_n1.json 2> gpurun_out/bench16_n1.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench16_n1.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['path']['kernel_ms_per_step'], d['clocks'])
PY
