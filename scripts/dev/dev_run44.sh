#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tests/tools/bench_mixed.py --steps 1 --only m --variants x:x:1 16:x:0 16:x:1 20:x:0 20:x:1 24:x:1 x:3e-3:0 x:3e-3:1 20:3e-3:1 x:1e-2:0 x:1e-2:1 20:1e-2:1 x:3e-2:1 > gpurun_out/bench_mixed_44.json 2> gpurun_out/bench_mixed_44.err; echo "bench rc=$?"
python - <<'P'
import json
j=json.loads([l for l in open('gpurun_out/bench_mixed_44.json') if l.startswith('{')][0])
print(j['m']['ms_per_solve'], j['m']['iterations'])
for v in j['m']['variants']: print(v)
P
tail -3 gpurun_out/bench_mixed_44.err
