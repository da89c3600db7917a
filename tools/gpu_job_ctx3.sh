#!/usr/bin/env bash
tag=${1:-ctx3}
mkdir -p gpurun_out
export CM_CONTEXTS=1000
timeout 600 python tools/bench_context.py > gpurun_out/${tag}_plain.jsonl 2> gpurun_out/${tag}_plain.log &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 3 -c 1 -o gpurun_out/${tag}_ctxscan -f python tools/bench_context.py > gpurun_out/${tag}_ncu.log 2>&1
echo "rc=$?"
