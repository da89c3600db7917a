set -x
timeout 900 python -m pytest tests/test_gpu_maintenance.py -m gpu -x -q > gpurun_out/r17_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r17_pytest.log
