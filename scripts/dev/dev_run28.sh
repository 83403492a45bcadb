#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1_28.json 2> gpurun_out/bench_n1_28.err; echo "bench rc=$?"
python -c "import json;j=json.loads([l for l in open('gpurun_out/bench_n1_28.json') if l.startswith('{')][0]);print(j['value'], j['ms_per_step'], j['e2e'], j['roofline']['frac'])"
