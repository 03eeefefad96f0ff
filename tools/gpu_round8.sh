#!/bin/bash
# 8-GPU validation of the scaling run the driver does at round end
mkdir -p gpurun_out
nvidia-smi -L | wc -l; free -g | head -2; nproc
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench8_n$N.json 2> gpurun_out/bench8_n$N.err; echo "bench n$N rc=$?"; grep -vE "^\*|OMP_NUM|^$" gpurun_out/bench8_n$N.err | tail -5; grep "^{" gpurun_out/bench8_n$N.json | cut -c1-2200
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>/dev/null | grep "^{" | cut -c1-600
