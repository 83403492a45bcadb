#!/usr/bin/env bash
# dev-only: build timing-ablation variants (a number) or the timeline-trace variant (the word 'trace', scripts/dev_tc_trace.py) of the tcgen05 product (TFQ_TC_ABLATE bit mask, see spmm_tc.cu) as
# tfqmrgpu_b200/lib/ablate/libtfQMRgpu_<mask>.so; time them on the GPU with
#   for m in ...; do TFQMRGPU_LIB=tfqmrgpu_b200/lib/ablate/libtfQMRgpu_$m.so python scripts/dev_spmm_time.py; done
set -euo pipefail
cd "$(dirname "$0")/../tfqmrgpu_b200/csrc"
make -j8 >/dev/null
mkdir -p ../lib/ablate
for m in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -diag-suppress 128 -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ \
       --expt-relaxed-constexpr $( [ "$m" = trace ] && echo -DTFQ_TC_TRACE || echo -DTFQ_TC_ABLATE=$m ) -c spmm_tc.cu -o ../lib/ablate/spmm_tc_$m.o
  objs=$(ls ../lib/obj/*.o | grep -v spmm_tc.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o ../lib/ablate/libtfQMRgpu_$m.so $objs ../lib/ablate/spmm_tc_$m.o \
       -L/usr/local/cuda/lib64 -lcurand -Xlinker -rpath,/usr/local/cuda/lib64
  echo "built lib/ablate/libtfQMRgpu_$m.so"
done
