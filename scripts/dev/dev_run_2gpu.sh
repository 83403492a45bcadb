set -x
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_multi_gpu.py "tests/test_gpu_parity.py::test_two_devices_in_one_process" -m gpu -q > gpurun_out/pytest_2gpu.log 2>&1; tail -25 gpurun_out/pytest_2gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 2500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
TFQMRGPU_BENCH_NO_PIPELINE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 1 --warmup 1 --precision z --ncols 16 --sigma 1 --tol 1e-9 --strong-rhs 0 > gpurun_out/bench_cfg4_n2.json 2> gpurun_out/bench_cfg4_n2.err; tail -c 2500 gpurun_out/bench_cfg4_n2.json; tail -5 gpurun_out/bench_cfg4_n2.err
