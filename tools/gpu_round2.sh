#!/bin/bash
mkdir -p gpurun_out
run() { echo "--- $*"; timeout 90 ./build/gemm_test "$@"; echo "exit=$?"; }
{
run 1 256 256 64 0
run 1 1000 768 1152 1
run 1 1000 768 1152 2
run 1 333 136 96 0
run 1 333 136 96 2
run 1 100 72 64 2
run 2 1000 768 1152 0
run 2 1000 776 1152 2
run 2 86400 768 768 2 20
run 2 86400 768 768 0 20
run 2 86400 2304 768 0 20
run 2 86400 3072 768 1 20
run 2 86400 768 3072 2 20
run 2 86400 3584 768 2 20
run 1 86400 3072 768 1 20
run 2 148992 9216 3584 0 5
} 2>&1 | tee gpurun_out/gemm_tests2.log | grep -E "^---|verify|time|exit=[1-9]"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench2.err; cat gpurun_out/bench2.json
