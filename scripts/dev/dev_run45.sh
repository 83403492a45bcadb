#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mixed or initial_guess or resident or early_freeze or max_iterations or fd_problem" > gpurun_out/pytest_45.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_45.log
timeout 400 python tests/tools/bench_mixed.py --steps 2 > gpurun_out/bench_mixed_45.json 2> gpurun_out/bench_mixed_45.err; echo "bench rc=$?"
grep "^{" gpurun_out/bench_mixed_45.json | cut -c 1-1800
tail -3 gpurun_out/bench_mixed_45.err
