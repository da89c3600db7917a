#!/usr/bin/env bash
# What the round-1 numbers under profiles/ were produced with. Run on a B200 box from the repo root, e.g.
#   gpurun --timeout 3000 -- 'bash tools/gpu_validate.sh'
# Outputs land in gpurun_out/. ncu runs only after the same command has exited 0 without ncu (B200_PROFILING.md).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.log; echo "bench exact rc=$?"
python bench.py --workload hnsw --steps 10 --warmup 3 > gpurun_out/bench_hnsw.json 2> gpurun_out/bench_hnsw.log; echo "bench hnsw rc=$?"
python bench.py --workload hnsw --build gpu --steps 10 --warmup 3 > gpurun_out/bench_hnsw_gpubuild.json 2> gpurun_out/bench_hnsw_gpubuild.log
python tools/diag_join.py > gpurun_out/join.log 2>&1
# launch list (shares) and full captures of the dominant kernels
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain1.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_exact.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 1 -c 1 -o gpurun_out/scan_tc \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
# (separate gpurun call: one ncu/--set full pass per call)
#   python bench.py --workload hnsw --efs 128 --steps 1 --warmup 1 --no-cpu-baseline && ncu --set full ... -k regex:k_hnsw_search ...
# read here with: python tools/ncu_key_metrics.py gpurun_out/scan_tc.ncu-rep ; python tools/launch_shares.py gpurun_out/launches_exact.csv
