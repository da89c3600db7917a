set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scan_tc.py -m gpu -x -q -k "gpu_assisted or 400000 or 60000" > gpurun_out/r16_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r16_pytest.log
timeout 900 python bench.py --workload hnsw --build gpu --steps 5 --warmup 2 > gpurun_out/r16_hnsw_gpubuild.json 2> gpurun_out/r16_hnsw_gpubuild.log; echo "hnsw gpu-build rc=$?"; grep "\[bench\]" gpurun_out/r16_hnsw_gpubuild.log | tail -8
