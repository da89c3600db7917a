#!/usr/bin/env bash
# 1M bench: one batch in flight vs two (streams / caller threads), K1 with 5 vs 6 stages.
tag=${1:-job}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "async_calls" > gpurun_out/${tag}_pytest_new.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest_new.log
for cfg in "1 1 6" "1 1 5" "2 2 6" "2 2 5"; do
  set -- $cfg
  CM_SCAN_STAGES=$3 timeout 600 python bench.py --steps 20 --warmup 5 --streams $1 --callers $2 --no-hnsw --no-cpu-baseline > gpurun_out/${tag}_bench_s$1_c$2_st$3.json 2> gpurun_out/${tag}_bench_s$1_c$2_st$3.log
  echo "streams=$1 callers=$2 stages=$3 rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.log; echo "bench default rc=$?"
python - "$tag" <<'PY'
import json, sys, glob
tag = sys.argv[1]
for f in sorted(glob.glob("gpurun_out/%s_bench*.json" % tag)):
    try:
        b = json.load(open(f))
        print(f.split("/")[-1], "value %.0f e2e %.0f ms %.3f K1 %.3f frac %.3f" % (b["value"], b["e2e"]["value"], b["ms_per_step"], b["roofline"]["kernel_ms_avg"], b["roofline"]["frac"]), (b.get("hnsw") or {}).get("frontier"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
