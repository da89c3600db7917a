#!/usr/bin/env bash
tag=${1:-job7}
mkdir -p gpurun_out
SECONDS=0; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.log; echo "bench rc=$?"
echo "bench wall ${SECONDS}s"
grep -E "source:|search_in_context|exact:" gpurun_out/${tag}_bench.log | cut -c1-400
python - "$tag" <<'PY'
import json, sys
b = json.load(open("gpurun_out/%s_bench.json" % sys.argv[1]))
print("value %.0f e2e %.0f ms %.3f K1 %.3f frac %.3f" % (b["value"], b["e2e"]["value"], b["ms_per_step"], b["roofline"]["kernel_ms_avg"], b["roofline"]["frac"]))
print("frontier", b["hnsw"]["frontier"])
print("ctx", b.get("search_in_context"))
print("src", b["config"]["source"])
PY
