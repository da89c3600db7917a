#!/usr/bin/env bash
tag=${1:-job}
mkdir -p gpurun_out
CM_BUILD_TIMING=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest_gpu.log
CM_BUILD_TIMING=1 timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.log; echo "reference rc=$?"
grep build_graph_gpu gpurun_out/${tag}_bench_reference.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.log; echo "bench rc=$?"
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
b = json.load(open("gpurun_out/%s_bench.json" % tag))
print("value %.0f e2e %.0f ms %.3f K1 %.3f frac %.3f" % (b["value"], b["e2e"]["value"], b["ms_per_step"], b["roofline"]["kernel_ms_avg"], b["roofline"]["frac"]))
print("frontier", b["hnsw"]["frontier"])
r = json.load(open("gpurun_out/%s_bench_reference.json" % tag))
print("ref %.0f ef %s build %.1fs" % (r["value"], r["config"]["ef"], r["config"]["graph_build_s"]))
PY
