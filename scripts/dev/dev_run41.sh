#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "resident or fd_problem or early_freeze or julia" > gpurun_out/pytest_41.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_41.log
timeout 300 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_cfg2_41.json 2> /dev/null
python -c "import json;j=json.loads([l for l in open('gpurun_out/bench_cfg2_41.json') if l.startswith('{')][0]);print('cfg2', j['value'], j['unit'], j.get('gpu_launches'))"
timeout 900 python bench.py --config 5 --no-cpu > gpurun_out/bench_cfg5_41.json 2> gpurun_out/bench_cfg5_41.err; echo "cfg5 rc=$?"
python -c "
import json;a=json.loads([l for l in open('gpurun_out/bench_cfg5_41.json') if l.startswith('{')][0])
print(a['value'], a['config']['not_converged'])
for r in a['config']['rows']:
    if r['lm']<=8 and r['rhs']<=64 and r['prec']=='c': print(r['lm'], r['ln'], r['rhs'], r['prec'], round(r['solve_ms'],2), r['iterations'])
"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
