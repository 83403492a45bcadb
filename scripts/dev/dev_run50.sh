#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu50.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu50.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
