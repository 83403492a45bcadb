set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu10.log 2>&1; tail -8 gpurun_out/pytest_gpu10.log
timeout 600 python bench.py --config 1 --steps 20 --warmup 3 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; tail -c 1500 gpurun_out/bench_cfg1.json; tail -3 gpurun_out/bench_cfg1.err
python bench.py --steps 2 --warmup 1 --no-cpu --strong-rhs 0 > gpurun_out/plain_r02.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:tfq -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu --strong-rhs 0 > gpurun_out/ncu_r02.log 2>&1
tail -c 600 gpurun_out/plain_r02.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke10.log 2>&1; tail -2 gpurun_out/smoke10.log
