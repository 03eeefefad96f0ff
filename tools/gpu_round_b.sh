#!/bin/bash
# SVA after the K|V fold, the host-streaming test, and the SURVEY 8.0/8d variant workloads.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sva.py tests/test_gpu_parity.py::test_compress_host_streams_the_same_bits -q -s 2>&1 | grep -E "^sva|masked|host-streamed|passed|failed|^E" | head -30
timeout 600 python tools/bench_sva.py 2>&1 | tail -1 | tee gpurun_out/bench_sva_v2.json
for W in literal_d1152_mlp segment_kv_d1152; do
  timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --cpu-sample-rows 24 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  echo "$W rc=$?"; tail -c 1500 gpurun_out/bench_$W.err | tail -5
done
timeout 600 python bench.py --workload eval64x600 --num-query 64 --segments 1200 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_eval64x600_k64.json 2> gpurun_out/bench_eval64x600_k64.err
echo "k64 rc=$?"
python - <<'PY'
import json
for n in ("literal_d1152_mlp","segment_kv_d1152","eval64x600_k64"):
    try:
        d=json.loads([l for l in open(f'gpurun_out/bench_{n}.json') if l.startswith('{')][-1])
        print(n, round(d['value']), 'video-s/s', round(d['ms_per_step'],1),'ms', 'kv', round(d['roofline']['achieved'] or 0), 'TF path', round(d['path']['algorithmic_tflops']), d['path']['kernel_ms_per_step'], (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(n, 'failed', e)
PY
