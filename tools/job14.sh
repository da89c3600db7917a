set -x
timeout 900 python -m pytest tests/test_gpu_context.py tests/test_golden.py -m gpu -x -q > gpurun_out/r14_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r14_pytest.log
