#!/bin/bash
mkdir -p gpurun_out
{
nvidia-smi topo -m
echo "--- numa"; for n in /sys/devices/system/node/node*; do echo "$n: $(cat $n/cpulist)  $(grep MemTotal $n/meminfo)"; done
echo "--- cpus allowed: $(python -c 'import os;print(len(os.sched_getaffinity(0)))')"
for B in 1 0; do
PROBE_BIND=$B timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tools/h2d_probe.py 2>&1 | grep -v "^\*\|OMP_NUM"
done
} > gpurun_out/h2d_probe.txt 2>&1
tail -5 gpurun_out/h2d_probe.txt
