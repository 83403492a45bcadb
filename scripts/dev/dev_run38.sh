#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu38.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu38.log
TFQMRGPU_BENCH_TIMELINE=1 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1_38.json 2> gpurun_out/bench_n1_38.err; echo "bench rc=$?"
grep "upload" gpurun_out/bench_n1_38.err | head -6
python -c "import json;j=json.loads([l for l in open('gpurun_out/bench_n1_38.json') if l.startswith('{')][0]);print(j['value'], j['ms_per_step'], j['e2e'], j['roofline']['frac'], j['clocks'])"
