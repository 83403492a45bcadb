#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu30.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu30.log
timeout 300 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/plain_cfg2_30.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:tfq -c 200 --csv --log-file gpurun_out/launches_fd_r02.csv python bench.py --config 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_fd30.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:resident_solve -c 1 -s 3 -o gpurun_out/prof_r02_resident python bench.py --config 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_res30.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/prof_r02_resident.ncu-rep
