#!/usr/bin/env bash
tag=${1:-late4}
mkdir -p gpurun_out
CM_RS_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_scan_tc.py -m gpu -x -q -s > gpurun_out/${tag}_pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -2 gpurun_out/${tag}_pytest_tc.log; grep -h "rescore\]" gpurun_out/${tag}_pytest_tc.log | head -5
for e in 0 1; do
  CM_SCAN_EPI2=$e timeout 300 python tools/diag_scan_floor.py >> gpurun_out/${tag}_floor.jsonl 2>> gpurun_out/${tag}_floor.log
  CM_ROWS=1000000 CM_SCAN_EPI2=$e timeout 300 python tools/diag_scan_floor.py >> gpurun_out/${tag}_floor.jsonl 2>> gpurun_out/${tag}_floor.log
  CM_ROWS=1000000 CM_DIM=64 CM_K=10 CM_SCAN_EPI2=$e timeout 300 python tools/diag_scan_floor.py >> gpurun_out/${tag}_floor.jsonl 2>> gpurun_out/${tag}_floor.log
  CM_ROWS=1000000 CM_DIM=128 CM_K=10 CM_SCAN_EPI2=$e timeout 300 python tools/diag_scan_floor.py >> gpurun_out/${tag}_floor.jsonl 2>> gpurun_out/${tag}_floor.log
done
cat gpurun_out/${tag}_floor.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 3 -c 1 -f -o gpurun_out/${tag}_scan_cfg4 env CM_SCAN_EPI2=0 python tools/diag_scan_floor.py > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
