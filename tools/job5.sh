set -x
timeout 900 python bench.py --workload hnsw --steps 5 --warmup 2 > gpurun_out/r5_hnsw.json 2> gpurun_out/r5_hnsw.log; echo "hnsw bench rc=$?"; cat gpurun_out/r5_hnsw.json; grep "bench" gpurun_out/r5_hnsw.log | tail -12
python bench.py --workload hnsw --rows 200000 --queries 2000 --efs 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r5_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_hnsw_search -s 1 -c 1 -o gpurun_out/r5_hnsw python bench.py --workload hnsw --rows 200000 --queries 2000 --efs 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r5_ncu.log 2>&1
echo "ncu rc=$?"
