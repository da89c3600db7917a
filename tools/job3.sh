set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3_ncu1.log 2>&1
echo "launches rc=$?"
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 1 -c 1 -o gpurun_out/r3_scan_tc2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3_ncu2.log 2>&1
echo "ncu full rc=$?"
