set -x
timeout 900 python -m pytest "tests/test_reference_callers.py::test_config3_full_size_vs_reference_cuda_build" -m gpu -q > gpurun_out/pytest_cfg3ref.log 2>&1; tail -5 gpurun_out/pytest_cfg3ref.log
timeout 600 python bench.py --config 1 --steps 20 --warmup 3 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; tail -c 1800 gpurun_out/bench_cfg1.json; tail -3 gpurun_out/bench_cfg1.err
timeout 600 python bench.py --config 2 --steps 20 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 1500 gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
timeout 1200 python bench.py --config 5 --sweep-rhs 64 --no-cpu > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; tail -c 600 gpurun_out/bench_cfg5.json; tail -3 gpurun_out/bench_cfg5.err
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -c 3000 gpurun_out/bench_r02.json; tail -3 gpurun_out/bench_r02.err
python bench.py --steps 2 --warmup 1 --no-cpu --strong-rhs 0 > gpurun_out/plain_r02.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu --strong-rhs 0 > gpurun_out/ncu_r02.log 2>&1
python scripts/dev_spmm_time.py > gpurun_out/plain_spmm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmm_tc16p -s 3 -c 2 -o gpurun_out/prof_r02_spmm_tc16p python scripts/dev_spmm_time.py > gpurun_out/ncu_full_r02.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
