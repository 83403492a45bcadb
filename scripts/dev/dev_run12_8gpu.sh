#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $T tests/tools/dev_exchange_probe.py gpurun_out/exchange_probe_n8.txt 5 > gpurun_out/exchange_probe_n8.log 2>&1
echo "probe rc=$?"
cat gpurun_out/exchange_probe_n8.txt
nvidia-smi --query-gpu=index,clocks.sm,power.draw,power.limit,temperature.gpu --format=csv > gpurun_out/smi_n8.txt
