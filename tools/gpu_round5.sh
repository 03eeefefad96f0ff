#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench5_n2.json 2> gpurun_out/bench5_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/bench5_n2.err; cat gpurun_out/bench5_n2.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdc_gemm_kernel -s 62 -c 1 -o gpurun_out/prof_kv_gemm_v2 \
   python bench.py --steps 1 --warmup 1 --segments 600 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu kv rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdc_attention_kernel -s 18 -c 2 -o gpurun_out/prof_attention_v2 \
   python bench.py --steps 1 --warmup 1 --segments 600 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; echo "ncu attn rc=$?"
