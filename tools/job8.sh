set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r8_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r8_pytest_gpu.log
timeout 900 python bench.py --workload hnsw --steps 5 --warmup 2 > gpurun_out/r8_hnsw.json 2> gpurun_out/r8_hnsw.log; echo "hnsw bench rc=$?"; grep "bench" gpurun_out/r8_hnsw.log | tail -12
