#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -x -q 2>&1 | tail -3
# sanitizer pass over the small-geometry parity cases (every kernel class runs)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "small_ or freq1 or batching or query_set" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "small_text or freq1" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/sanitizer_racecheck.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench12_n1.json 2> gpurun_out/bench12_n1.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench12_n1.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
