#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/tools/bench_inprocess_multi.py 8 3 > gpurun_out/inprocess_n8.json 2> gpurun_out/inprocess_n8.err; echo "inprocess rc=$?"
tail -3 gpurun_out/inprocess_n8.err; cat gpurun_out/inprocess_n8.json | cut -c1-900
TFQMRGPU_NUM_GPUS=8 timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -k "not torchrun and not nccl" > gpurun_out/pytest_multi_n8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_multi_n8.log
