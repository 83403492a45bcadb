#!/bin/bash
# round-2 scaling run on 8 GPUs of one box: weak + strong legs of config 3, config 4 as written (1024 right-hand sides over 8 GPUs)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
NCCL_DEBUG=WARN timeout 700 $T bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "n8 rc=$?"
timeout 500 $T bench.py --gpus 8 --config 4 --steps 2 --warmup 3 > gpurun_out/bench_cfg4_n8.json 2> gpurun_out/bench_cfg4_n8.err
echo "cfg4 n8 rc=$?"
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
tail -c 600 gpurun_out/bench_n8.json
