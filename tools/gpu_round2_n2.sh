#!/bin/bash
# Round 2 multi-GPU evidence (gpurun --gpus N): per-device state, exchange correctness, strong scaling.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_round2_n2.sh 2'
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
{
echo "## two engines on two devices in one process"
timeout 300 python tools/two_device_check.py 2>&1 | tail -5
echo "## multicast gather == NCCL all-gather (world $N)"
timeout 300 $TR tools/mcast_check.py 2>&1 | grep -E "mcast_check|Error|error" | head -5
echo "## sharded compress_video == single GPU (world $N)"
timeout 300 $TR tools/dist_check.py 2>&1 | grep -E "dist_check|Error|error" | head -5
} > gpurun_out/r02_multigpu_checks_n$N.txt 2>&1
cat gpurun_out/r02_multigpu_checks_n$N.txt
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n$N.err; cat gpurun_out/r02_bench_n$N.json
