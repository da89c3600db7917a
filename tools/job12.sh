set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r12_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r12_pytest_gpu.log
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/r12_bench.json 2> gpurun_out/r12_bench.log; echo "bench rc=$?"; cat gpurun_out/r12_bench.json; tail -3 gpurun_out/r12_bench.log
timeout 600 python bench.py --steps 40 --warmup 3 --data uniform > gpurun_out/r12_bench_uniform.json 2> gpurun_out/r12_bench_uniform.log; echo "bench uniform rc=$?"; cat gpurun_out/r12_bench_uniform.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r12_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r12_launches_exact.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r12_ncu1.log 2>&1
echo "launches rc=$?"
