#!/bin/bash
mkdir -p gpurun_out
python tools/audio_dbg.py > gpurun_out/audio_dbg.log 2>&1; tail -5 gpurun_out/audio_dbg.log
for D in 0 1; do echo "=== TDC_GEMM_DEBUG=$D (1 = no epilogue math/stores)"; for cfg in "2 86400 3072 768 1" "2 86400 768 768 2" "2 86400 2304 768 0" "1 86400 768 768 2" "2 86400 768 3072 2"; do TDC_GEMM_DEBUG=$D ./build/gemm_test $cfg 20 | grep -E "time" | sed "s/^/$cfg : /"; done; done 2>&1 | tee gpurun_out/gemm_noepi.log
