#!/usr/bin/env bash
tag=${1:-sweep}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_scan_tc.py -m gpu -x -q > gpurun_out/${tag}_pytest_tc.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_tc.log
timeout 600 python tools/sweep_scan.py > gpurun_out/${tag}_sweep.jsonl 2> gpurun_out/${tag}_sweep.log; echo "sweep rc=$?"; cat gpurun_out/${tag}_sweep.jsonl; tail -3 gpurun_out/${tag}_sweep.log
