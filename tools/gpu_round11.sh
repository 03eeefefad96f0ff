#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/bench_segment.py 2>&1 | tail -2 | tee gpurun_out/bench_segment.json
timeout 900 python bench.py --workload cfg2_llama3b --steps 10 --warmup 3 > gpurun_out/bench11_cfg2.json 2> gpurun_out/bench11_cfg2.err; echo "bench cfg2 rc=$?"; tail -3 gpurun_out/bench11_cfg2.err; cut -c1-1800 gpurun_out/bench11_cfg2.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench11_n1.json 2> gpurun_out/bench11_n1.err; echo "bench n1 rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench11_n1.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['path']['kernel_ms_per_step'], d['clocks'], d['roofline']['achieved'], d['e2e']['ms_per_step'])
PY
