set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r20_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r20_pytest_gpu.log
