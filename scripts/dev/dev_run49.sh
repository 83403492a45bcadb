#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -k "mixed or initial_guess or preconditioner or c_example or julia" > gpurun_out/pytest_49.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_49.log
./examples/_build/mixed_precision 64 16 2 1e-10
timeout 400 python tests/tools/bench_mixed.py --steps 2 > gpurun_out/bench_mixed_49.json 2> gpurun_out/bench_mixed_49.err; echo "bench rc=$?"
grep "^{" gpurun_out/bench_mixed_49.json | cut -c 1-1800
tail -3 gpurun_out/bench_mixed_49.err
