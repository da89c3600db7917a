set -x
timeout 600 python -m pytest tests/test_gpu_scan_tc.py -m gpu -x -q > gpurun_out/r4_pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -15 gpurun_out/r4_pytest_tc.log
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.log; echo "bench rc=$?"; cat gpurun_out/r4_bench.json; tail -5 gpurun_out/r4_bench.log
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r4_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 1 -c 1 -o gpurun_out/r4_scan_tc3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r4_ncu2.log 2>&1
echo "ncu full rc=$?"
