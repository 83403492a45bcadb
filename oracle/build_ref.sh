#!/usr/bin/env bash
# TEST INFRASTRUCTURE (oracle). Builds the UNMODIFIED reference from /root/reference into oracle/_ref/.
#   cpu : reference CPU path (-DHAS_NO_CUDA, the reference's own fallback; SURVEY.md §8c recipe)
#   gen : the reference's FD example generator (example/tfqmrgpu_generate_FD_example.cxx)
#   gpu : reference CUDA kernels compiled for sm_100 (same-box GPU baseline; optional)
#   benchcpu : the reference's bench harness on the reference CPU path (bench_tfqmrgpu.cu -DHAS_NO_CUDA + libtfqmr_ref_cpu.so)
#             -> _ref/bench_tfqmrgpu_cpu, plus _ref/libalign256.so (LD_PRELOAD: 256-byte aligned malloc, which the CPU path's
#             workspace allocator silently assumes); used to check that the reference's own readers accept files we write
#   callers : the reference's OWN callers, unmodified, linked against OUR libtfQMRgpu.so (drop-in evidence):
#             example/tfqmrgpu_C_example.c -> _ref/c_example_ours ; source/bench_tfqmrgpu.cu -> _ref/bench_tfqmrgpu_ours
#             (and bench_tfqmrgpu_ref linked against _ref/libtfqmr_ref_gpu.so for same-box comparisons)
# Sources are compiled where they lie; nothing is copied. Outputs go to oracle/_ref/ only (git-ignored).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${TFQMR_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
[ -d "$REF/tfQMRgpu" ] || { echo "reference not present at $REF - nothing built"; exit 0; }
mkdir -p "$OUT"
INC="-I$REF/tfQMRgpu/include -I$REF/tfQMRgpu/source -I$REF/third_party/rapidxml-1.13"
# plain -O1/-O2 miscompile the reference's out-of-bounds flattened indexing (UB); these flags agree with -O0
SAFE="-O3 -fno-tree-dce -fno-aggressive-loop-optimizations -fno-strict-aliasing"
what="${1:-all}"
if [ "$what" = all ] || [ "$what" = cpu ]; then
  g++ -std=c++14 $SAFE -fopenmp -fPIC -shared -DHAS_NO_CUDA -include "$HERE/ref_shim.h" $INC \
      -x c++ "$REF/tfQMRgpu/source/tfqmrgpu.cu" -x none "$HERE/ref_harness.cpp" -DHARNESS_NO_CUDA \
      -o "$OUT/libtfqmr_ref_cpu.so"
  echo "built $OUT/libtfqmr_ref_cpu.so"
fi
if [ "$what" = all ] || [ "$what" = gen ]; then
  g++ -std=c++14 $SAFE -D__MAIN__ -include cstdint -include cmath -include cstdlib $INC \
      "$REF/example/tfqmrgpu_generate_FD_example.cxx" -o "$OUT/generate_FD_example"
  echo "built $OUT/generate_FD_example"
fi
if [ "$what" = all ] || [ "$what" = gpu ]; then
  if command -v nvcc >/dev/null; then
    nvcc -std=c++14 -O2 -arch=sm_100 -Xcompiler=-fopenmp,-fPIC,-fno-tree-dce,-fno-aggressive-loop-optimizations \
        -include cstdint $INC -shared "$REF/tfQMRgpu/source/tfqmrgpu.cu" "$HERE/ref_harness.cpp" \
        -o "$OUT/libtfqmr_ref_gpu.so" -lcurand
    echo "built $OUT/libtfqmr_ref_gpu.so"
  fi
fi
if [ "$what" = all ] || [ "$what" = benchcpu ]; then
  if [ -f "$OUT/libtfqmr_ref_cpu.so" ]; then
    g++ -std=c++14 $SAFE -fopenmp -DHAS_NO_CUDA -include "$HERE/ref_shim.h" $INC \
        -x c++ "$REF/tfQMRgpu/source/bench_tfqmrgpu.cu" -o "$OUT/bench_tfqmrgpu_cpu" \
        -L"$OUT" -ltfqmr_ref_cpu -Wl,-rpath,'$ORIGIN'
    printf '#include <stdlib.h>\nvoid *malloc(size_t n) { void *p = 0; return posix_memalign(&p, 256, n ? n : 1) ? 0 : p; }\n' \
      | gcc -O2 -fPIC -shared -x c - -o "$OUT/libalign256.so"
    echo "built $OUT/bench_tfqmrgpu_cpu"
  fi
fi
if [ "$what" = all ] || [ "$what" = callers ]; then
  OURS="$HERE/../tfqmrgpu_b200/lib"
  if [ -f "$OURS/libtfQMRgpu.so" ]; then
    gcc -O2 -DHAS_TFQMRGPU "$REF/example/tfqmrgpu_C_example.c" -o "$OUT/c_example_ours" \
        -L"$OURS" -ltfQMRgpu -lm -Wl,-rpath,'$ORIGIN/../../tfqmrgpu_b200/lib'
    echo "built $OUT/c_example_ours"
    if command -v nvcc >/dev/null; then
      nvcc -std=c++14 -O2 -arch=sm_100 -Xcompiler=-fopenmp,-fno-tree-dce,-fno-aggressive-loop-optimizations \
          -include cstdint $INC "$REF/tfQMRgpu/source/bench_tfqmrgpu.cu" -o "$OUT/bench_tfqmrgpu_ours" \
          -L"$OURS" -ltfQMRgpu -Xlinker -rpath,'$ORIGIN/../../tfqmrgpu_b200/lib'
      echo "built $OUT/bench_tfqmrgpu_ours"
      if [ -f "$OUT/libtfqmr_ref_gpu.so" ]; then
        nvcc -std=c++14 -O2 -arch=sm_100 -Xcompiler=-fopenmp,-fno-tree-dce,-fno-aggressive-loop-optimizations \
            -include cstdint $INC "$REF/tfQMRgpu/source/bench_tfqmrgpu.cu" -o "$OUT/bench_tfqmrgpu_ref" \
            -L"$OUT" -ltfqmr_ref_gpu -lcurand -Xlinker -rpath,'$ORIGIN'
        echo "built $OUT/bench_tfqmrgpu_ref"
      fi
    fi
  fi
fi
