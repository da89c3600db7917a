set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r18_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r18_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r18_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r18_smoke.log
timeout 600 python bench.py > gpurun_out/r18_bench_exact.json 2> gpurun_out/r18_bench_exact.log; echo "bench rc=$?"; cat gpurun_out/r18_bench_exact.json | cut -c1-700
timeout 900 python bench.py --workload hnsw --steps 10 --warmup 3 > gpurun_out/r18_bench_hnsw.json 2> gpurun_out/r18_bench_hnsw.log; echo "hnsw rc=$?"; grep "\[bench\]" gpurun_out/r18_bench_hnsw.log | tail -6 | cut -c1-600
timeout 600 python tools/diag_join.py > gpurun_out/r18_join.log 2>&1; echo "join rc=$?"; grep "^rep" gpurun_out/r18_join.log
python bench.py --workload hnsw --efs 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r18_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_hnsw_search|rescore_kernel" -s 1 -c 2 -o gpurun_out/r18_hnsw_final python bench.py --workload hnsw --efs 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r18_ncu.log 2>&1
echo "ncu rc=$?"
