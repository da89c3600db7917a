#!/usr/bin/env bash
tag=${1:-late5}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scan_tc.py tests/test_gpu_context.py tests/test_gpu_maintenance.py -m gpu -x -q > gpurun_out/${tag}_pytest_tc.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest_tc.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gpu_assisted_build or device_select or exact" > gpurun_out/${tag}_pytest_build.log 2>&1; echo "pytest build rc=$?"; tail -2 gpurun_out/${tag}_pytest_build.log
for e in 0 1; do
  CM_SCAN_EPI2=$e timeout 300 python tools/diag_scan_floor.py >> gpurun_out/${tag}_floor.jsonl 2>> gpurun_out/${tag}_floor.log
  CM_ROWS=1000000 CM_SCAN_EPI2=$e timeout 300 python tools/diag_scan_floor.py >> gpurun_out/${tag}_floor.jsonl 2>> gpurun_out/${tag}_floor.log
  CM_ROWS=1000000 CM_DIM=64 CM_K=10 CM_SCAN_EPI2=$e timeout 300 python tools/diag_scan_floor.py >> gpurun_out/${tag}_floor.jsonl 2>> gpurun_out/${tag}_floor.log
done
cat gpurun_out/${tag}_floor.jsonl
timeout 300 python bench.py --no-hnsw --no-cpu-baseline --steps 20 --warmup 5 --source matrix > gpurun_out/${tag}_full.json 2> gpurun_out/${tag}_full.log; echo "1M rc=$?"
python - <<PY
import json
b = json.load(open("gpurun_out/${tag}_full.json"))
fl = b["config"].get("batches_in_flight", {})
print("1M: value %.0f ms %.3f K1 %.3f e2e %.0f" % (b["value"], b["ms_per_step"], b["roofline"]["kernel_ms_avg"], b["e2e"]["value"]), fl.get("device_qps_one_in_flight"), fl.get("device_qps_two_in_flight"))
PY
