set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k hnsw > gpurun_out/r9_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r9_pytest_gpu.log
timeout 900 python tools/diag_hnsw_rows.py > gpurun_out/r9_rows.log 2>&1; echo "rc=$?"; grep -v "^\[bench\]" gpurun_out/r9_rows.log | tail -30
