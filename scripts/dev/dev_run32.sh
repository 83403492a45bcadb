#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -q -x -k "early_freeze or resident or breakdown or set_devices or solve_vs_oracle" > gpurun_out/pytest_32.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_32.log
timeout 900 python bench.py --config 5 --no-cpu > gpurun_out/bench_cfg5_32.json 2> gpurun_out/bench_cfg5_32.err; echo "cfg5 rc=$?"
python -c "
import json;a=json.loads([l for l in open('gpurun_out/bench_cfg5_32.json') if l.startswith('{')][0])
print(a['value'], a['config']['not_converged'])
for r in a['config']['rows']:
    if 'with_early_freeze' in r: print(r['lm'], r['ln'], r['rhs'], r['status'], r['iterations'], r['residual'], r['with_early_freeze'])
"
