#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu13.log
timeout 300 python bench.py --config 2 --steps 20 --warmup 3 > gpurun_out/bench_cfg2_13.json 2> gpurun_out/bench_cfg2_13.err; echo "cfg2 rc=$?"
timeout 600 python bench.py --config 5 --no-cpu > gpurun_out/bench_cfg5_13.json 2> gpurun_out/bench_cfg5_13.err; echo "cfg5 rc=$?"
