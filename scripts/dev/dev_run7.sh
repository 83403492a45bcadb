set -x
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_multi7.log 2>&1; tail -30 gpurun_out/pytest_multi7.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu7.log 2>&1; tail -15 gpurun_out/pytest_gpu7.log
