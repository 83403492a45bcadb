set -x
timeout 900 python tests/tools/dev_tc16_check.py > gpurun_out/tc16_check2.log 2>&1; echo "rc=$?" >> gpurun_out/tc16_check2.log
grep -E "FAIL|PASS|rc=|Error|error" gpurun_out/tc16_check2.log | head
for c in 2 4 8 16; do TFQMRGPU_TC_CHAIN=$c timeout 300 python tests/tools/dev_tc16_acc.py >> gpurun_out/tc16_acc.log 2>&1; done
for c in 2 4 8 16 32; do TFQMRGPU_TC_CHAIN=$c TFQMRGPU_DEV_SKIP_XOP=1 timeout 300 python scripts/dev_spmm_time.py >> gpurun_out/tc16_time2.log 2>&1; done
for m in 1 4 7; do TFQMRGPU_DEV_SKIP_XOP=1 TFQMRGPU_LIB=tfqmrgpu_b200/lib/ablate/libtfQMRgpu16_$m.so timeout 300 python scripts/dev_spmm_time.py >> gpurun_out/tc16_time2.log 2>&1; done
TFQMRGPU_LIB=$PWD/tfqmrgpu_b200/lib/ablate/libtfQMRgpu16_trace.so TFQMRGPU_DEV_SKIP_XOP=1 timeout 300 python scripts/dev_tc16_trace.py > gpurun_out/tc16_trace.log 2>&1
TFQMRGPU_TC_CHAIN=4 TFQMRGPU_LIB=$PWD/tfqmrgpu_b200/lib/ablate/libtfQMRgpu16_trace.so TFQMRGPU_DEV_SKIP_XOP=1 timeout 300 python scripts/dev_tc16_trace.py >> gpurun_out/tc16_trace.log 2>&1
cat gpurun_out/tc16_time2.log gpurun_out/tc16_trace.log gpurun_out/tc16_acc.log
