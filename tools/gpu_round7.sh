#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/mcast_check.py 2>&1 | grep -vE "^\*|OMP_NUM|^$" | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/dist_check.py 2>&1 | grep -E "dist_check|Error" | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench7_n2.json 2> gpurun_out/bench7_n2.err; echo "bench n2 rc=$?"; grep -vE "^\*|OMP_NUM|^$" gpurun_out/bench7_n2.err | tail -5; grep "^{" gpurun_out/bench7_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-multicast > gpurun_out/bench7_n2_nccl.json 2> gpurun_out/bench7_n2_nccl.err; echo "bench n2 nccl rc=$?"; grep "^{" gpurun_out/bench7_n2_nccl.json | cut -c1-400
