#!/usr/bin/env bash
tag=${1:-job}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest_gpu.log
CM_BUILD_TIMING=1 timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.log; echo "reference rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.log; echo "bench rc=$?"
timeout 600 python tools/bench_context.py > gpurun_out/${tag}_context.jsonl 2> gpurun_out/${tag}_context.log; echo "context rc=$?"; cut -c1-420 gpurun_out/${tag}_context.jsonl
# full ncu captures of the two dominant kernels (each only after the same command ran clean without ncu)
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-hnsw > gpurun_out/${tag}_plain2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 1 -c 1 -o gpurun_out/${tag}_scan_tc \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-hnsw > gpurun_out/${tag}_ncu2.log 2>&1
echo "ncu scan rc=$?"
timeout 300 python bench.py --workload hnsw --efs 32 --ef 32 --hnsw-steps 3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_plain3.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hnsw_search -s 2 -c 1 -o gpurun_out/${tag}_hnsw_ef32 \
      python bench.py --workload hnsw --efs 32 --ef 32 --hnsw-steps 3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu3.log 2>&1
echo "ncu hnsw rc=$?"
