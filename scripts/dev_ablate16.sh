#!/usr/bin/env bash
# dev-only: timing-ablation variants of the fp16-pair tcgen05 product (TFQ_TC16_ABLATE bit mask, see spmm_tc16.cu) as
# tfqmrgpu_b200/lib/ablate/libtfQMRgpu16_<mask>.so; time them on the GPU with
#   for m in ...; do TFQMRGPU_LIB=tfqmrgpu_b200/lib/ablate/libtfQMRgpu16_$m.so python scripts/dev_spmm_time.py; done
set -euo pipefail
cd "$(dirname "$0")/../tfqmrgpu_b200/csrc"
make -j8 >/dev/null
mkdir -p ../lib/ablate
for m in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -diag-suppress 128 -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ \
       --expt-relaxed-constexpr $( [ "$m" = trace ] && echo -DTFQ_TC16_TRACE || echo -DTFQ_TC16_ABLATE=$m ) -c spmm_tc16.cu -o ../lib/ablate/spmm_tc16_$m.o
  objs=$(ls ../lib/obj/*.o | grep -v spmm_tc16.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o ../lib/ablate/libtfQMRgpu16_$m.so $objs ../lib/ablate/spmm_tc16_$m.o \
       -L/usr/local/cuda/lib64 -lcurand -Xlinker -rpath,/usr/local/cuda/lib64
  echo "built lib/ablate/libtfQMRgpu16_$m.so"
done
