#!/usr/bin/env bash
# Round-2 (late) experiment: histogram threshold (Ladder) for k > 32 in K1's epilogue, K2 rows per round at a
# 1/8 shard (the per-rank work of the 8-GPU run), launch shares of that step.
tag=${1:-lad}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan_tc.py -m gpu -x -q > gpurun_out/${tag}_pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -2 gpurun_out/${tag}_pytest_tc.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gpu_assisted_build or device_select" > gpurun_out/${tag}_pytest_build.log 2>&1; echo "pytest build rc=$?"; tail -2 gpurun_out/${tag}_pytest_build.log
cfg4="--rows 4000000 --dim 32 --k 100 --data uniform --no-hnsw --no-cpu-baseline --steps 10 --warmup 3 --source matrix"
for lad in 1 0; do
  CM_SCAN_LADDER=$lad timeout 300 python bench.py $cfg4 > gpurun_out/${tag}_cfg4_ladder${lad}.json 2> gpurun_out/${tag}_cfg4_ladder${lad}.log; echo "cfg4 ladder=$lad rc=$?"
done
shard="--rows 125000 --no-hnsw --no-cpu-baseline --steps 60 --warmup 5 --source matrix"
for rows in 8 6 5 4; do
  CM_RS_ROWS=$rows timeout 300 python bench.py $shard > gpurun_out/${tag}_shard_rs${rows}.json 2> gpurun_out/${tag}_shard_rs${rows}.log; echo "shard rs=$rows rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_shard.csv \
    python bench.py --rows 125000 --no-hnsw --no-cpu-baseline --steps 3 --warmup 2 --streams 1 --callers 1 --source matrix > gpurun_out/${tag}_ncu_shard.log 2>&1
echo "ncu shard rc=$?"
CM_BUILD_TIMING=1 timeout 400 python bench.py --workload hnsw --efs 28 --ef 28 --hnsw-steps 3 --steps 3 --warmup 1 --no-cpu-baseline --no-graph-cache --source matrix > gpurun_out/${tag}_build.json 2> gpurun_out/${tag}_build.log; echo "build rc=$?"
grep -i "knn\|lists\|total\|graph:" gpurun_out/${tag}_build.log | head -20
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_cfg4_*.json") + glob.glob("gpurun_out/${tag}_shard_*.json")):
    try:
        b = json.load(open(f))
        fl = b["config"].get("batches_in_flight", {})
        print(f, "value %.0f ms %.3f K1 %.3f e2e %.0f" % (b["value"], b["ms_per_step"], b["roofline"]["kernel_ms_avg"], b["e2e"]["value"]),
              fl.get("device_qps_one_in_flight"), fl.get("device_qps_two_in_flight"), b["config"].get("rescored_per_query_e2e"), b["config"].get("fallback_queries_e2e"))
    except Exception as e:
        print(f, "unreadable", e)
PY
