#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "resident or fd_problem or julia or breakdown" > gpurun_out/pytest_42.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_42.log
timeout 300 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_cfg2_42.json 2> /dev/null
python -c "import json;j=json.loads([l for l in open('gpurun_out/bench_cfg2_42.json') if l.startswith('{')][0]);print('cfg2', j['value'], j['unit'], j.get('gpu_launches'))"
TFQMRGPU_RESIDENT_TRACE=1 timeout 300 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu 2>&1 | grep "# resident" | tail -1
