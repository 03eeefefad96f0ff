#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sva.py -q -s 2>&1 | grep -E "^sva|masked|passed|failed|^E" | head -30
timeout 600 python tools/bench_sva.py 2>&1 | tail -1 | tee gpurun_out/bench_sva.json
