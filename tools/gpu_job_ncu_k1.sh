#!/usr/bin/env bash
# ncu --set full of K1 (new default: resident query blocks) and of the all-streamed setting, one capture each
tag=${1:-k1}
mkdir -p gpurun_out
cmd="python bench.py --steps 1 --warmup 1 --streams 1 --callers 1 --no-cpu-baseline --no-hnsw"
timeout 300 $cmd > gpurun_out/${tag}_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 1 -c 1 -f -o gpurun_out/${tag}_scan_tc_qres7 $cmd > gpurun_out/${tag}_ncu_a.log 2>&1
echo "ncu qres7 rc=$?"
export CM_SCAN_QRES=0 CM_SCAN_SLOTS=10
timeout 300 $cmd > gpurun_out/${tag}_plain_b.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 1 -c 1 -f -o gpurun_out/${tag}_scan_tc_qres0 $cmd > gpurun_out/${tag}_ncu_b.log 2>&1
echo "ncu qres0 rc=$?"
