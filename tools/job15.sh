set -x
timeout 900 python -m pytest tests/test_gpu_scan_tc.py -m gpu -x -q > gpurun_out/r15_pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -3 gpurun_out/r15_pytest_tc.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r15_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/r15_memcheck_smoke.log
