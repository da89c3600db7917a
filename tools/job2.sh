set -x
timeout 600 python -m pytest tests/test_gpu_scan_tc.py -m gpu -x -q > gpurun_out/r2_pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -15 gpurun_out/r2_pytest_tc.log
timeout 600 python tools/diag_tc.py > gpurun_out/r2_diag_tc.log 2>&1; echo "diag rc=$?"; tail -30 gpurun_out/r2_diag_tc.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.log; echo "bench rc=$?"; cat gpurun_out/r2_bench.json; tail -5 gpurun_out/r2_bench.log
