#!/usr/bin/env bash
# Full GPU suite on the current build, then an A/B of the 1M x 768 headline: the library built from the previous commit
# (tools/_ab/libchronomind_b200_head.so) against the current one, alternating runs in one call on one box.
tag=${1:-late6}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest_gpu.log
lib=chrono-mind_b200/libchronomind_b200.so
cp $lib /tmp/lib_new.so
run() {  # name
  timeout 300 python bench.py --no-hnsw --no-cpu-baseline --steps 20 --warmup 5 --source matrix > gpurun_out/${tag}_$1.json 2> gpurun_out/${tag}_$1.log
  python - "$1" gpurun_out/${tag}_$1.json <<'PY'
import json, sys
b = json.load(open(sys.argv[2]))
f = b["config"]["batches_in_flight"]
print("%-8s one %.0f (%.3f ms, K1 %.3f) two %.0f  e2e %s" % (sys.argv[1], f["device_qps_one_in_flight"], f["ms_per_step_one_in_flight"], b["roofline"]["kernel_ms_avg"], f["device_qps_two_in_flight"], f["e2e_qps_by_caller_threads"]))
PY
}
for rep in 1 2; do
  cp tools/_ab/libchronomind_b200_head.so $lib; run head$rep
  cp /tmp/lib_new.so $lib; run new$rep
done
cp /tmp/lib_new.so $lib
for e in 0 1; do
  CM_ROWS=1000000 CM_DIM=64 CM_K=10 CM_SCAN_EPI2=$e timeout 300 python tools/diag_scan_floor.py >> gpurun_out/${tag}_floor.jsonl 2>> gpurun_out/${tag}_floor.log
  CM_SCAN_EPI2=$e timeout 300 python tools/diag_scan_floor.py >> gpurun_out/${tag}_floor.jsonl 2>> gpurun_out/${tag}_floor.log
done
cat gpurun_out/${tag}_floor.jsonl
