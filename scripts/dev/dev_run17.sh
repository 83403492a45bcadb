#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "resident or fd_problem or solve_every_block_size or max_iterations or breakdown or rhs_trivial" > gpurun_out/pytest_res17.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_res17.log
timeout 300 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_cfg2_res17.json 2> gpurun_out/bench_cfg2_res17.err; echo "cfg2 rc=$?"
python -c "import json;j=json.loads([l for l in open('gpurun_out/bench_cfg2_res17.json') if l.startswith('{')][0]);print('resident', j['value'], j['unit'], j['config'].get('iterations'), j.get('gpu_launches'))"
TFQMRGPU_RESIDENT_TRACE=1 timeout 300 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu 2>&1 | grep "# resident" | tail -3
TFQMRGPU_RESIDENT_A=0 TFQMRGPU_RESIDENT_TRACE=1 timeout 300 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu 2>&1 | grep "# resident" | tail -2
