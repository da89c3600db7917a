#!/usr/bin/env bash
tag=${1:-late3}
mkdir -p gpurun_out
CM_RS_DEBUG=1 timeout 300 python -m pytest tests/test_gpu_scan_tc.py -m gpu -x -q -s -k ladder > gpurun_out/${tag}_dbg_e1.log 2>&1; echo "dbg e1 rc=$?"
CM_RS_DEBUG=1 CM_SCAN_EPI2=0 timeout 300 python -m pytest tests/test_gpu_scan_tc.py -m gpu -x -q -s -k ladder > gpurun_out/${tag}_dbg_e0.log 2>&1; echo "dbg e0 rc=$?"
grep -h "rescore\]" gpurun_out/${tag}_dbg_e1.log | head -5; grep -h "rescore\]" gpurun_out/${tag}_dbg_e0.log | head -5
for e in 0 1; do
  CM_SCAN_EPI2=$e timeout 300 python tools/diag_scan_floor.py >> gpurun_out/${tag}_floor.jsonl 2>> gpurun_out/${tag}_floor.log
  CM_ROWS=1000000 CM_SCAN_DIAG=1 CM_SCAN_EPI2=$e timeout 300 python tools/diag_scan_floor.py >> gpurun_out/${tag}_floor.jsonl 2>> gpurun_out/${tag}_floor.log
  CM_ROWS=1000000 CM_SCAN_EPI2=$e timeout 300 python tools/diag_scan_floor.py >> gpurun_out/${tag}_floor.jsonl 2>> gpurun_out/${tag}_floor.log
done
cat gpurun_out/${tag}_floor.jsonl
