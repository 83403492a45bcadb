#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu24.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu24.log
timeout 300 python bench.py --config 2 --steps 20 --warmup 3 > gpurun_out/bench_cfg2_24.json 2> gpurun_out/bench_cfg2_24.err; echo "cfg2 rc=$?"
timeout 900 python bench.py --config 5 --no-cpu > gpurun_out/bench_cfg5_24.json 2> gpurun_out/bench_cfg5_24.err; echo "cfg5 rc=$?"
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "resident_solver_matches and (fd or 8-c-3 or 5-c-2)" > gpurun_out/racecheck_resident.log 2>&1; echo "racecheck rc=$?"
tail -8 gpurun_out/racecheck_resident.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "resident" > gpurun_out/memcheck_resident.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/memcheck_resident.log
