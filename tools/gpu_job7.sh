#!/usr/bin/env bash
tag=${1:-job7}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_snapshot_scale.py tests/test_gpu_context.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
/usr/bin/time -v timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.log; echo "bench rc=$?"
grep -E "Elapsed|Maximum resident" gpurun_out/${tag}_bench.log
grep -E "source:|search_in_context|exact:" gpurun_out/${tag}_bench.log | cut -c1-400
python - "$tag" <<'PY'
import json, sys
b = json.load(open("gpurun_out/%s_bench.json" % sys.argv[1]))
print("value %.0f e2e %.0f ms %.3f K1 %.3f frac %.3f" % (b["value"], b["e2e"]["value"], b["ms_per_step"], b["roofline"]["kernel_ms_avg"], b["roofline"]["frac"]))
print("frontier", b["hnsw"]["frontier"])
print("ctx", b.get("search_in_context"))
print("src", b["config"]["source"])
PY
