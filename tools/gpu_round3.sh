#!/bin/bash
mkdir -p gpurun_out
run() { echo "--- $*"; timeout 90 ./build/gemm_test "$@"; echo "exit=$?"; }
{
run 2 86400 3072 768 1 20
run 1 86400 3072 768 1 20
run 2 1000 768 1152 1
} 2>&1 | grep -E "^---|verify|time|exit=[1-9]"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"; tail -3 gpurun_out/bench3.err; cat gpurun_out/bench3.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01.csv \
   python bench.py --steps 1 --warmup 1 --segments 600 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdc_gemm_kernel -s 62 -c 1 -o gpurun_out/prof_kv_gemm \
   python bench.py --steps 1 --warmup 1 --segments 600 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu kv rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdc_attention_kernel -s 18 -c 2 -o gpurun_out/prof_attention \
   python bench.py --steps 1 --warmup 1 --segments 600 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; echo "ncu attn rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdc_gemm_kernel -s 63 -c 6 -o gpurun_out/prof_query_gemm \
   python bench.py --steps 1 --warmup 1 --segments 600 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1; echo "ncu qgemm rc=$?"
ls -la gpurun_out
