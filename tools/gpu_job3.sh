#!/usr/bin/env bash
tag=${1:-job}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python tools/bench_context.py > gpurun_out/${tag}_context.jsonl 2> gpurun_out/${tag}_context.log; echo "context rc=$?"; cat gpurun_out/${tag}_context.jsonl | cut -c1-700
timeout 900 python tools/diag_hnsw_rows.py > gpurun_out/${tag}_diag_hnsw.log 2>&1; echo "diag rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.log; echo "bench rc=$?"
