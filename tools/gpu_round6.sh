#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
run() { echo "--- $*"; timeout 90 ./build/gemm_test "$@"; echo "exit=$?"; }
{ run 2 148992 9216 3584 0 5; run 2 86400 2304 768 0 20; } 2>&1 | grep -E "^---|verify|time|exit=[1-9]"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/symm_probe.py > gpurun_out/symm_probe.log 2>&1; echo "probe rc=$?"
grep -E "NCCL all_gather|symm_mem|via |NVLS|P2P|SHM" gpurun_out/symm_probe.log | head -30
timeout 900 ncu --set full --clock-control none -k regex:tdc_gemm_kernel -s 62 -c 1 -o gpurun_out/prof_kv_gemm_v3 \
   python bench.py --steps 1 --warmup 1 --segments 600 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu kv rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench6_n1.json 2> gpurun_out/bench6_n1.err; echo "bench n1 rc=$?"; tail -3 gpurun_out/bench6_n1.err; cat gpurun_out/bench6_n1.json
