#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "small_text or freq1" > gpurun_out/sanitizer_racecheck2.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck2.log
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "small_text or freq1" > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/sanitizer_synccheck.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench13_n1.json 2> gpurun_out/bench13_n1.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench13_n1.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['path']['kernel_ms_per_step'], d['clocks'])
PY
