#!/bin/bash
# Round-2 ncu / sanitizer evidence (1 GPU):  gpurun --timeout 1500 -- 'bash tools/gpu_round2_evidence.sh'
mkdir -p gpurun_out
python tools/kv_gemm_shape.py 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tdc_gemm_kernel -s 4 -c 1 -o gpurun_out/r02_prof_kv_gemm_folded python tools/kv_gemm_shape.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tdc_gemm_ln_kernel -s 30 -c 1 -o gpurun_out/r02_prof_gemm_ln_k3072 python tools/bench_gemm_ln.py 1800 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:tdc_attention_kernel -s 40 -c 4 -o gpurun_out/r02_prof_attention python bench.py --steps 1 --warmup 0 --segments 1800 --no-e2e --no-cpu-baseline --no-qformer-only --unfolded-steps 0 --parity-rows 0 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
{
echo "## memcheck: tests/test_gpu_frames.py -k 'reference_driver or errors'"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_frames.py -q -k "reference_driver or errors" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|=========.*at " | head -20
echo "## memcheck: tests/test_gpu_parity.py -k 'small_text or small_notext' (fused GEMM + LayerNorm, cluster DSMEM)"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -k "small_text or small_notext" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|=========.*at " | head -20
echo "## memcheck: tests/test_gpu_sva.py -k 'two_query or bilinear'; tests/test_gpu_modules.py -k 'speech_qformer_module or adapt_segment'"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_sva.py tests/test_gpu_modules.py -q -k "two_query or bilinear or speech_qformer_module or adapt_segment" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|=========.*at " | head -20
echo "## racecheck: tests/test_gpu_parity.py -k small_notext"
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -k "small_notext" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | sort | uniq -c | sort -rn | head -12
} > gpurun_out/r02_sanitizer.txt 2>&1
cat gpurun_out/r02_sanitizer.txt
