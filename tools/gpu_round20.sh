#!/bin/bash
mkdir -p gpurun_out
run() { echo "--- MC=$TDC_GEMM_MC $*"; timeout 60 ./build/gemm_test "$@"; echo "exit=$?"; }
for MC in 2 1; do export TDC_GEMM_MC=$MC
{ run 2 700 768 1152 2; run 2 86400 768 768 2 20; run 2 86400 2304 768 0 20; run 2 86400 3072 768 1 20; run 2 86400 768 3072 2 20; run 2 148992 9216 3584 0 5; } 2>&1 | grep -E "^---|verify|time|exit=[1-9]|mismatch|timed out|error"
done | tee gpurun_out/gemm_mc2.log
