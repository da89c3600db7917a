#!/usr/bin/env bash
# A/B of K1 settings under the bench's own conditions (device-resident batches back to back): alternate runs in one call
tag=${1:-ab}
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-hnsw --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.log
  python - "$name" gpurun_out/${tag}_${name}.json <<'PY'
import json, sys
b = json.load(open(sys.argv[2]))
f = b["config"]["batches_in_flight"]
print("%-10s one %.0f (%.3f ms, K1 %.3f) two %.0f  e2e %s" % (sys.argv[1], f["device_qps_one_in_flight"], f["ms_per_step_one_in_flight"], b["roofline"]["kernel_ms_avg"], f["device_qps_two_in_flight"], f["e2e_qps_by_caller_threads"]))
PY
}
for rep in 1 2; do
  run old${rep} CM_SCAN_QRES=0 CM_SCAN_SLOTS=10
  run new${rep} CM_NOP=1
done
