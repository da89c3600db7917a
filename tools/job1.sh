set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r1_pytest_gpu.log
python tools/diag_tc.py > gpurun_out/r1_diag_tc.log 2>&1; echo "diag rc=$?"; tail -30 gpurun_out/r1_diag_tc.log
python bench.py --engine 2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_tc1.csv python bench.py --engine 2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_ncu1.log 2>&1
echo "launches rc=$?"
python bench.py --engine 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 2 -c 2 -o gpurun_out/r1_scan_tc1 python bench.py --engine 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1_ncu2.log 2>&1
echo "ncu full rc=$?"
