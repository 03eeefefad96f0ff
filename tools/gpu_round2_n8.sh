#!/bin/bash
# Round 2, one 8-GPU box:  gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_round2_n8.sh'
#  1. exchange correctness at world 8 (multicast gather == NCCL, sharded == single GPU)
#  2. the default bench line at N = 8: weak scaling + exchange_check + the strong-scaling record (ONE 1-hour video
#     over 8 GPUs) + end to end from pinned host memory
#  3. BASELINE config 5: 64 x 10-minute videos over 8 GPUs (4800 video-seconds per GPU), K = 16 and K = 64
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
{
echo "## multicast gather == NCCL all-gather (world $N)"
timeout 300 $TR tools/mcast_check.py 2>&1 | grep -E "mcast_check|Error|error" | head -5
echo "## sharded compress_video == single GPU (world $N)"
timeout 300 $TR tools/dist_check.py 2>&1 | grep -E "dist_check|Error|error" | head -5
} > gpurun_out/r02_multigpu_checks_n$N.txt 2>&1
cat gpurun_out/r02_multigpu_checks_n$N.txt
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 2> gpurun_out/r02_bench_n$N.err | grep "^{" > gpurun_out/r02_bench_n$N.json
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n$N.err; cut -c1-1500 gpurun_out/r02_bench_n$N.json
timeout 600 $TR bench.py --gpus $N --workload eval64x600 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --strong-steps 2 --parity-rows 12 2> gpurun_out/r02_cfg5_k16_n$N.err | grep "^{" > gpurun_out/r02_cfg5_k16_n$N.json
echo "cfg5 K=16 rc=$?"; tail -2 gpurun_out/r02_cfg5_k16_n$N.err; cut -c1-600 gpurun_out/r02_cfg5_k16_n$N.json
timeout 600 $TR bench.py --gpus $N --workload eval64x600 --num-query 64 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --strong-steps 2 --parity-rows 12 2> gpurun_out/r02_cfg5_k64_n$N.err | grep "^{" > gpurun_out/r02_cfg5_k64_n$N.json
echo "cfg5 K=64 rc=$?"; tail -2 gpurun_out/r02_cfg5_k64_n$N.err; cut -c1-600 gpurun_out/r02_cfg5_k64_n$N.json
