#!/bin/bash
# dev script: run under gpurun.  Each case is its own process + timeout so a wedged
# kernel cannot take the rest of the run (or the box) with it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
run() { echo "--- $*"; timeout 90 ./build/gemm_test "$@"; echo "exit=$?"; }
{
run 1 256 256 64 0
run 1 128 128 128 0
run 1 1000 768 1152 0
run 1 1000 768 1152 1
run 1 1000 768 1152 2
run 1 333 136 96 0
run 1 32768 768 768 2 20
run 1 32768 3072 768 1 20
run 1 32768 768 3072 2 20
run 1 148992 9216 3072 0 5
run 2 512 256 64 0
run 2 1000 768 1152 0
run 2 1000 768 1152 2
run 2 32768 3072 768 1 20
run 2 148992 9216 3072 0 5
} 2>&1 | tee gpurun_out/gemm_tests.log
