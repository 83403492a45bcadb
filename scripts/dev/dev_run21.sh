#!/bin/bash
for abl in 0 1 2 3; do
TFQMRGPU_RESIDENT_ABLATE=$abl TFQMRGPU_RESIDENT_TRACE=1 timeout 300 python -m tfqmrgpu_b200.bench_cli tfQMR tests/golden/FD_problem.xml z 3 41 2>&1 | grep "# resident" | tail -1 | sed "s/^/ablate $abl: /"
done
