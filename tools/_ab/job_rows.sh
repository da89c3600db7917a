mkdir -p gpurun_out
rm -f gpurun_out/rows_pb.jsonl
for pb in 0 1 0 1; do
  echo "{\"park_best\": $pb}" >> gpurun_out/rows_pb.jsonl
  CM_SCAN_PARK_BEST=$pb CM_SIZES=62500,125000,1000000 timeout 600 python tools/diag_scan_rows.py >> gpurun_out/rows_pb.jsonl 2>> gpurun_out/rows_pb.log
done
cat gpurun_out/rows_pb.jsonl
timeout 600 python -m pytest tests/test_gpu_scan_tc.py tests/test_gpu_context.py -m gpu -x -q > gpurun_out/rows_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/rows_pytest.log
