#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/bench_sva.py 2>&1 | tail -1 | tee gpurun_out/bench_sva_v3.json
