#!/bin/bash
# Everything the driver checks at round end, in one gpurun call (1 GPU):
#   gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh'
# Outputs land in gpurun_out/ (scratch); copy what should be judged into profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cut -c1-400 gpurun_out/bench_reference.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
if [ "$1" == "--sanitize" ] || [ "$2" == "--sanitize" ]; then
  # memcheck over the kernels added after profiles/r01_sanitizer.txt: attention key mask, cyclic-residual LayerNorm,
  # residual add (SVA), segmentation, pooling, text mode with the last layer's text tokens skipped
  {
  echo "## memcheck: tests/test_gpu_sva.py -k 'small or kv_mask'"
  timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_sva.py -q -k "small or kv_mask" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|=========.*at " | head -20
  echo "## memcheck: tests/test_gpu_modules.py -k 'adapt_segment or avg_pool or audio or compressor'"
  timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_modules.py -q -k "adapt_segment or avg_pool or audio or compressor" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|=========.*at " | head -20
  echo "## memcheck: tests/test_gpu_parity.py -k 'small_text or host_streams'"
  timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -k "small_text or host_streams" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|=========.*at " | head -20
  } > gpurun_out/sanitizer_v2.txt 2>&1
  cat gpurun_out/sanitizer_v2.txt
fi
if [ "$1" == "--ncu" ]; then
  B="python bench.py --steps 1 --warmup 1 --segments 600 --no-e2e --no-cpu-baseline"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $B > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdc_gemm_kernel -s 62 -c 1 -o gpurun_out/prof_kv_gemm $B > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdc_attention_kernel -s 18 -c 2 -o gpurun_out/prof_attention $B > /dev/null 2>&1
  ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
fi
