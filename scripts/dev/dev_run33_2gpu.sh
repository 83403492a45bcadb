#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 700 $T bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_33.json 2> gpurun_out/bench_n2_33.err; echo "n2 rc=$?"
python -c "import json;j=json.loads([l for l in open('gpurun_out/bench_n2_33.json') if l.startswith('{')][0]);print(j['value'], j['ms_per_step'], j['e2e'], j.get('strong'))"
tail -3 gpurun_out/bench_n2_33.err
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/pytest_multi_33.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_multi_33.log
