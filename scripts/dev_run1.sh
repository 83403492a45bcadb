set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 300 python tests/tools/dev_tc16_check.py quick > gpurun_out/tc16_quick.log 2>&1; echo "rc=$?" >> gpurun_out/tc16_quick.log
tail -5 gpurun_out/tc16_quick.log
if grep -q "rc=0" gpurun_out/tc16_quick.log; then
  timeout 900 python tests/tools/dev_tc16_check.py > gpurun_out/tc16_check.log 2>&1; echo "rc=$?" >> gpurun_out/tc16_check.log
  for t in 1 2; do TFQMRGPU_TENSOR=$t timeout 300 python scripts/dev_spmm_time.py >> gpurun_out/tc16_time.log 2>&1; done
  TFQMRGPU_DEV_SKIP_XOP=1 timeout 300 python scripts/dev_spmm_time.py >> gpurun_out/tc16_time.log 2>&1
  for m in 1 3 4 7 8 16; do TFQMRGPU_DEV_SKIP_XOP=1 TFQMRGPU_LIB=tfqmrgpu_b200/lib/ablate/libtfQMRgpu16_$m.so timeout 300 python scripts/dev_spmm_time.py >> gpurun_out/tc16_time.log 2>&1; done
  for c in 4 8 32; do TFQMRGPU_TC_CHAIN=$c TFQMRGPU_DEV_SKIP_XOP=1 timeout 300 python scripts/dev_spmm_time.py >> gpurun_out/tc16_time.log 2>&1; done
  cat gpurun_out/tc16_time.log
fi
