#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tdc_gemm -s 2 -c 1 -o gpurun_out/prof_gemm_n768_f32 ./build/gemm_test 2 86400 768 768 2 3 > /dev/null 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tdc_gemm -s 2 -c 1 -o gpurun_out/prof_gemm_gelu ./build/gemm_test 2 86400 3072 768 1 3 > /dev/null 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep | tail -3
