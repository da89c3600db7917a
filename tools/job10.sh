set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r10_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r10_launches_exact.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r10_ncu1.log 2>&1
echo "launches rc=$?"
python bench.py --workload hnsw --efs 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r10_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_hnsw_search -s 1 -c 1 -o gpurun_out/r10_hnsw_v3 python bench.py --workload hnsw --efs 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r10_ncu2.log 2>&1
echo "ncu rc=$?"
