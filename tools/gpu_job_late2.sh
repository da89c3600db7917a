#!/usr/bin/env bash
# Late round-2 validation of: K2 without block-wide sorts (rank counting, async query row), K1 large-k Ladder,
# second epilogue warp-quad for short rows, threshold prefetch. Full GPU suite, then A/B benches.
tag=${1:-late2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest_gpu.log
cfg4="--rows 4000000 --dim 32 --k 100 --data uniform --no-hnsw --no-cpu-baseline --steps 10 --warmup 3 --source matrix"
CM_SCAN_EPI2=1 CM_SCAN_LADDER=1 timeout 300 python bench.py $cfg4 > gpurun_out/${tag}_cfg4_e1l1.json 2> gpurun_out/${tag}_cfg4_e1l1.log; echo "cfg4 e1l1 rc=$?"
CM_SCAN_EPI2=0 CM_SCAN_LADDER=1 timeout 300 python bench.py $cfg4 > gpurun_out/${tag}_cfg4_e0l1.json 2> gpurun_out/${tag}_cfg4_e0l1.log; echo "cfg4 e0l1 rc=$?"
CM_SCAN_EPI2=1 CM_SCAN_LADDER=0 timeout 300 python bench.py $cfg4 > gpurun_out/${tag}_cfg4_e1l0.json 2> gpurun_out/${tag}_cfg4_e1l0.log; echo "cfg4 e1l0 rc=$?"
shard="--rows 125000 --no-hnsw --no-cpu-baseline --steps 60 --warmup 5 --source matrix"
timeout 300 python bench.py $shard > gpurun_out/${tag}_shard.json 2> gpurun_out/${tag}_shard.log; echo "shard rc=$?"
timeout 300 python bench.py --no-hnsw --no-cpu-baseline --steps 20 --warmup 5 --source matrix > gpurun_out/${tag}_full.json 2> gpurun_out/${tag}_full.log; echo "1M rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_shard.csv \
    python bench.py --rows 125000 --no-hnsw --no-cpu-baseline --steps 3 --warmup 2 --streams 1 --callers 1 --source matrix > gpurun_out/${tag}_ncu_shard.log 2>&1
echo "ncu shard rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        b = json.load(open(f))
        fl = b["config"].get("batches_in_flight", {})
        print(f, "value %.0f ms %.3f K1 %.3f e2e %.0f" % (b["value"], b["ms_per_step"], b["roofline"]["kernel_ms_avg"], b["e2e"]["value"]),
              fl.get("device_qps_one_in_flight"), fl.get("device_qps_two_in_flight"), b["config"].get("rescored_per_query_e2e"), b["config"].get("fallback_queries_e2e"))
    except Exception as e:
        print(f, "unreadable", e)
PY
python tools/launch_shares.py gpurun_out/${tag}_launches_shard.csv
