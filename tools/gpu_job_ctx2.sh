#!/usr/bin/env bash
tag=${1:-ctx2}
mkdir -p gpurun_out
export CM_CONTEXTS=1000
timeout 600 python tools/bench_context.py > gpurun_out/${tag}_plain.jsonl 2> gpurun_out/${tag}_plain.log &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv python tools/bench_context.py > gpurun_out/${tag}_ncu.log 2>&1
echo "rc=$?"
python tools/launch_shares.py gpurun_out/${tag}_launches.csv | head -30
