#!/usr/bin/env bash
# context-sorted search_in_context: parity tests, then the 1M x 768 numbers (CM_CONTEXT_UNSORTED=1 = the mask-only scan)
tag=${1:-ctx}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_context.py tests/test_gpu_maintenance.py -m gpu -x -q > gpurun_out/${tag}_pytest_ctx.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_ctx.log
timeout 600 python tools/bench_context.py > gpurun_out/${tag}_context_sorted.jsonl 2> gpurun_out/${tag}_context_sorted.log; echo "sorted rc=$?"; cat gpurun_out/${tag}_context_sorted.jsonl
if [ -n "$CM_AB" ]; then
CM_CONTEXT_UNSORTED=1 timeout 600 python tools/bench_context.py > gpurun_out/${tag}_context_unsorted.jsonl 2> gpurun_out/${tag}_context_unsorted.log; echo "unsorted rc=$?"; cat gpurun_out/${tag}_context_unsorted.jsonl
fi
