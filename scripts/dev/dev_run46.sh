#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "preconditioner or user_defined or initial_guess" > gpurun_out/pytest_46.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_46.log
