#!/usr/bin/env bash
# What the round-2 numbers under profiles/ were produced with. Run on a B200 box from the repo root:
#   gpurun --timeout 3000 -- 'bash tools/gpu_validate.sh r2final'
# Outputs land in gpurun_out/<tag>_*. ncu runs only after the same command has exited 0 without ncu (B200_PROFILING.md).
tag=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
CM_BUILD_TIMING=1 timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.log; echo "reference rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.log; echo "bench rc=$?"
timeout 600 python tools/bench_context.py > gpurun_out/${tag}_context.jsonl 2> gpurun_out/${tag}_context.log; echo "context rc=$?"
# the per-rank work of the 8-GPU run (a 125k-row shard) on one GPU, and BASELINE configs[4]'s per-GPU shape
timeout 300 python bench.py --rows 125000 --no-hnsw --no-cpu-baseline --steps 60 --warmup 5 --source matrix > gpurun_out/${tag}_bench_shard125k.json 2> gpurun_out/${tag}_bench_shard125k.log; echo "shard rc=$?"
timeout 300 python bench.py --rows 4000000 --dim 32 --k 100 --data uniform --no-hnsw --no-cpu-baseline --steps 10 --warmup 3 --source matrix > gpurun_out/${tag}_bench_cfg4.json 2> gpurun_out/${tag}_bench_cfg4.log; echo "cfg4 rc=$?"
# launch list (shares) of the exact step, one batch in flight
timeout 300 python bench.py --steps 3 --warmup 2 --streams 1 --callers 1 --no-cpu-baseline --no-hnsw > gpurun_out/${tag}_plain1.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_exact.csv \
      python bench.py --steps 3 --warmup 2 --streams 1 --callers 1 --no-cpu-baseline --no-hnsw > gpurun_out/${tag}_ncu1.log 2>&1
echo "ncu launches rc=$?"
# full captures of the two dominant kernels
timeout 300 python bench.py --steps 1 --warmup 1 --streams 1 --callers 1 --no-cpu-baseline --no-hnsw > gpurun_out/${tag}_plain2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 1 -c 1 -o gpurun_out/${tag}_scan_tc \
      python bench.py --steps 1 --warmup 1 --streams 1 --callers 1 --no-cpu-baseline --no-hnsw > gpurun_out/${tag}_ncu2.log 2>&1
echo "ncu scan rc=$?"
timeout 300 python bench.py --workload hnsw --efs 28 --ef 28 --hnsw-steps 3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_plain3.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hnsw_search -s 2 -c 1 -o gpurun_out/${tag}_hnsw_ef28 \
      python bench.py --workload hnsw --efs 28 --ef 28 --hnsw-steps 3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu3.log 2>&1
echo "ncu hnsw rc=$?"
# read here with: python tools/ncu_key_metrics.py gpurun_out/<tag>_scan_tc.ncu-rep ; python tools/launch_shares.py gpurun_out/<tag>_launches_exact.csv
# compute-sanitizer (closed on this pool in round 1; tried again, bounded)
[ -n "$CM_SKIP_MEMCHECK" ] || timeout 240 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge_cases or all_rows_identical or known_answers" > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/${tag}_memcheck.log
