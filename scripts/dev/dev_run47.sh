#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu47.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu47.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_47.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_47.log
timeout 600 python bench.py > gpurun_out/bench_n1_47.json 2> gpurun_out/bench_n1_47.err; echo "bench rc=$?"
python - <<'P'
import json
j=json.loads([l for l in open('gpurun_out/bench_n1_47.json') if l.startswith('{')][0])
print('value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['ms_per_step'], 'clocks', j['clocks'])
P
timeout 300 python bench.py --config 2 --steps 20 --warmup 3 > gpurun_out/bench_cfg2_47.json 2> /dev/null; echo "cfg2 rc=$?"
python -c "import json;j=json.loads([l for l in open('gpurun_out/bench_cfg2_47.json') if l.startswith('{')][0]);print('cfg2', j['value'], j['unit'], j.get('reference_gpu',{}).get('value'), j.get('cpu_baseline',{}).get('value'))"
