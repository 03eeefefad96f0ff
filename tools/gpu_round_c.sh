#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "host_streams or golden" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c_n1.json 2>gpurun_out/bench_c_n1.err; echo rc=$?
timeout 600 python bench.py --steps 3 --warmup 3 --num-text 32 --no-cpu-baseline --no-e2e > gpurun_out/bench_c_text32.json 2>/dev/null; echo rc=$?
python - <<'PY'
import json
for n in ("c_n1","c_text32"):
    d=json.loads([l for l in open(f'gpurun_out/bench_{n}.json') if l.startswith('{')][-1])
    e=d.get('e2e') or {}
    print(n, round(d['value']), round(d['ms_per_step'],1),'ms kv', round(d['roofline']['achieved']), 'e2e', e.get('value') and round(e['value']), e.get('ms_per_step') and round(e['ms_per_step'],1), d['path']['kernel_ms_per_step'], d['gpu_launches'])
PY
