#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_callers.py -m gpu -q -x -k "resident or fd or FD or solve_every_block_size or max_iterations or breakdown or rhs_trivial or c_example or julia or fortran" > gpurun_out/pytest_res27.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_res27.log
timeout 900 python bench.py --config 5 --no-cpu > gpurun_out/bench_cfg5_27.json 2> gpurun_out/bench_cfg5_27.err; echo "cfg5 rc=$?"
