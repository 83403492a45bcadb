#!/bin/bash
mkdir -p gpurun_out
for t in 1 2; do
TFQMRGPU_RESIDENT_TILES_PER_SM=$t TFQMRGPU_RESIDENT_TRACE=1 timeout 300 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu 2>&1 | grep "# resident" | tail -1
done
