#!/usr/bin/env bash
# gpurun job: GPU tests of the exact path, bench (N=1), then the ncu launch list of the same bench command.
tag=${1:-job}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan_tc.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.log; echo "bench rc=$?"
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-hnsw > gpurun_out/${tag}_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_exact.csv \
      python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-hnsw > gpurun_out/${tag}_ncu1.log 2>&1
echo "ncu rc=$?"
python - "$tag" <<'PY'
import json, sys
b=json.load(open("gpurun_out/%s_bench.json" % sys.argv[1]))
print("value %.0f e2e %.0f ms %.3f K1 %.3f ms frac_burst %.3f launches %d" % (b["value"], b["e2e"]["value"], b["ms_per_step"], b["roofline"]["kernel_ms_avg"], b["roofline"]["frac"], b["gpu_launches"]))
print("frontier", b["hnsw"]["frontier"])
PY
