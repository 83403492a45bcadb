#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "resident or fd_problem or solve_every_block_size or max_iterations or breakdown" > gpurun_out/pytest_res15.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_res15.log
for c in 1 2 4; do
  TFQMRGPU_RESIDENT_CTAS=$c timeout 300 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_cfg2_res_c$c.json 2> gpurun_out/bench_cfg2_res_c$c.err; echo "cfg2 ctas=$c rc=$?"
  python -c "import json;j=json.loads([l for l in open('gpurun_out/bench_cfg2_res_c$c.json') if l.startswith('{')][0]);print($c, j['value'], j['unit'], j['config'].get('iterations'), j.get('gpu_launches'))"
done
TFQMRGPU_RESIDENT=0 timeout 300 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_cfg2_nores.json 2>&1
python -c "import json;j=json.loads([l for l in open('gpurun_out/bench_cfg2_nores.json') if l.startswith('{')][0]);print('off', j['value'], j['unit'], j['config'].get('iterations'), j.get('gpu_launches'))"
