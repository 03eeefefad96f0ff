#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench9_n1.json 2> gpurun_out/bench9_n1.err; echo "bench n1 rc=$?"; tail -3 gpurun_out/bench9_n1.err; cut -c1-2500 gpurun_out/bench9_n1.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdc_attention_kernel -s 18 -c 2 -o gpurun_out/prof_attention_v3 \
   python bench.py --steps 1 --warmup 1 --segments 600 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; echo "ncu attn rc=$?"
