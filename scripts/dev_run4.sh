set -x
timeout 900 python tests/tools/dev_tc16_check.py > gpurun_out/tc16_check5.log 2>&1; echo "rc=$?" >> gpurun_out/tc16_check5.log
grep -E "FAIL|PASS|rc=|Error|error|worst" gpurun_out/tc16_check5.log | head -20
TFQMRGPU_DEV_SKIP_XOP=1 timeout 300 python scripts/dev_spmm_time.py > gpurun_out/tc16_time5.log 2>&1
timeout 300 python scripts/dev_spmm_time.py >> gpurun_out/tc16_time5.log 2>&1
for m in 1 4 7; do TFQMRGPU_DEV_SKIP_XOP=1 TFQMRGPU_LIB=tfqmrgpu_b200/lib/ablate/libtfQMRgpu16_$m.so timeout 300 python scripts/dev_spmm_time.py >> gpurun_out/tc16_time5.log 2>&1; done
TFQMRGPU_LIB=$PWD/tfqmrgpu_b200/lib/ablate/libtfQMRgpu16_trace.so TFQMRGPU_DEV_SKIP_XOP=1 timeout 300 python scripts/dev_tc16_trace.py > gpurun_out/tc16_trace5.log 2>&1
cat gpurun_out/tc16_time5.log gpurun_out/tc16_trace5.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -c 1500 gpurun_out/bench5.json
