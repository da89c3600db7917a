#!/usr/bin/env bash
# ncu --set full (source-level) of K2 at a 1/8 shard and of K1's large-k (Ladder) epilogue on configs[4]; each only after
# the same command has exited 0 without ncu (the job before this one ran them).
tag=${1:-late}
mkdir -p gpurun_out
shard="python bench.py --rows 125000 --no-hnsw --no-cpu-baseline --steps 1 --warmup 1 --streams 1 --callers 1 --source matrix"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rescore_kernel -s 1 -c 1 -f -o gpurun_out/${tag}_rescore_shard $shard > gpurun_out/${tag}_ncu_a.log 2>&1
echo "ncu rescore rc=$?"
cfg4="python bench.py --rows 4000000 --dim 32 --k 100 --data uniform --no-hnsw --no-cpu-baseline --steps 1 --warmup 1 --streams 1 --callers 1 --source matrix"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 1 -c 1 -f -o gpurun_out/${tag}_scan_cfg4_ladder $cfg4 > gpurun_out/${tag}_ncu_b.log 2>&1
echo "ncu cfg4 rc=$?"
