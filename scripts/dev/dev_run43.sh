#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mixed or setmatrix_errors or max_iterations or julia" > gpurun_out/pytest_43.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_43.log
TFQMRGPU_VERBOSE=2 timeout 400 python tests/tools/bench_mixed.py --steps 2 > gpurun_out/bench_mixed_43.json 2> gpurun_out/bench_mixed_43.err; echo "bench rc=$?"
grep "mixed:" gpurun_out/bench_mixed_43.json | tail -12
grep "^{" gpurun_out/bench_mixed_43.json | cut -c 1-1800
tail -3 gpurun_out/bench_mixed_43.err
