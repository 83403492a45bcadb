#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/pytest_multi_48.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_multi_48.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29548"
timeout 400 $T bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --strong-rhs 0 > gpurun_out/bench_n2_48.json 2> gpurun_out/bench_n2_48.err; echo "n2 rc=$?"
python -c "import json;j=json.loads([l for l in open('gpurun_out/bench_n2_48.json') if l.startswith('{')][0]);print(j['value'], j['ms_per_step'], j['config']['iterations_per_solve_min_max_over_ranks'], j['e2e']['ms_per_step'], j['roofline']['frac'])"
tail -3 gpurun_out/bench_n2_48.err
