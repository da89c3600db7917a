#!/usr/bin/env bash
# gpurun --gpus N job: group tests, torchrun bench at N, --abi-group bench at N.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_job_multi.sh r2c 2'
tag=${1:-job}; N=${2:-2}; extra=${3:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_group.py -m gpu -x -q > gpurun_out/${tag}_pytest_group.log 2>&1; echo "pytest group rc=$?"; tail -3 gpurun_out/${tag}_pytest_group.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 $extra \
    > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.log; echo "torchrun bench rc=$?"
timeout 900 python bench.py --gpus $N --abi-group --steps 20 --warmup 5 --efs 32,64,128 > gpurun_out/${tag}_bench_group_${N}gpu.json 2> gpurun_out/${tag}_bench_group_${N}gpu.log; echo "group bench rc=$?"
python - "$tag" "$N" <<'PY'
import json, sys
tag, N = sys.argv[1], sys.argv[2]
for name in ("bench_%sgpu" % N, "bench_group_%sgpu" % N):
    try:
        b = json.load(open("gpurun_out/%s_%s.json" % (tag, name)))
        print(name, "value %.0f e2e %.0f ms %.3f" % (b["value"], b["e2e"]["value"], b["ms_per_step"]), (b.get("hnsw") or {}).get("frontier"))
    except Exception as e:
        print(name, "unreadable:", e)
PY
