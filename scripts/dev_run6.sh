set -x
for f in planar direct; do
  TFQMRGPU_TC_FORM=$f timeout 900 python tests/tools/dev_tc16_check.py > gpurun_out/tc16_check6_$f.log 2>&1; echo "rc=$?" >> gpurun_out/tc16_check6_$f.log
  grep -E "FAIL|PASS|rc=|Error|error|worst" gpurun_out/tc16_check6_$f.log | head -20
done
timeout 300 python tests/tools/dev_tc16_acc.py > gpurun_out/tc16_acc6.log 2>&1
TFQMRGPU_TC_FORM=planar timeout 300 python tests/tools/dev_tc16_acc.py >> gpurun_out/tc16_acc6.log 2>&1
for c in 4 8 16; do TFQMRGPU_TC_CHAIN=$c TFQMRGPU_DEV_SKIP_XOP=1 timeout 300 python scripts/dev_spmm_time.py >> gpurun_out/tc16_time6.log 2>&1; done
timeout 300 python scripts/dev_spmm_time.py >> gpurun_out/tc16_time6.log 2>&1
TFQMRGPU_TC_FORM=direct TFQMRGPU_DEV_SKIP_XOP=1 timeout 300 python scripts/dev_spmm_time.py >> gpurun_out/tc16_time6.log 2>&1
cat gpurun_out/tc16_time6.log gpurun_out/tc16_acc6.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; tail -15 gpurun_out/pytest_gpu6.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; tail -c 1500 gpurun_out/bench6.json
