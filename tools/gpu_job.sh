#!/usr/bin/env bash
# One gpurun job: GPU test suite, smoke, both bench arms. Outputs under gpurun_out/<tag>_*.
#   gpurun --timeout 1800 -- 'bash tools/gpu_job.sh r2a'
tag=${1:-job}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/${tag}_smi.txt 2>&1
nproc > gpurun_out/${tag}_nproc.txt; free -g >> gpurun_out/${tag}_nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.log; echo "bench reference rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.log; echo "bench rc=$?"
tail -c 1500 gpurun_out/${tag}_bench.json
