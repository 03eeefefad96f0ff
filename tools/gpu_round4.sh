#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
run() { echo "--- $*"; timeout 90 ./build/gemm_test "$@"; echo "exit=$?"; }
{ run 2 148992 9216 3584 0 5; run 2 86400 2304 768 0 20; run 2 1000 776 1152 2; } 2>&1 | grep -E "^---|verify|time|exit=[1-9]"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench4_n1.json 2> gpurun_out/bench4_n1.err; echo "bench n1 rc=$?"; tail -3 gpurun_out/bench4_n1.err; cat gpurun_out/bench4_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench4_n2.json 2> gpurun_out/bench4_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/bench4_n2.err; cat gpurun_out/bench4_n2.json
