#!/usr/bin/env bash
# BASELINE configs[3]: 10M x 768 sharded over N GPUs (exact scan; HNSW block only when asked for).
#   gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_job_10m.sh r2g N [extra bench args]'
tag=${1:-job}; N=${2:-1}; shift 2
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "async_calls or remove_then_sync or clustered" > gpurun_out/${tag}_pytest_new.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest_new.log
  timeout 1200 python bench.py --rows 10000000 --steps 10 --warmup 3 "$@" > gpurun_out/${tag}_bench_10M_1gpu.json 2> gpurun_out/${tag}_bench_10M_1gpu.log; echo "bench rc=$?"
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --rows 10000000 --steps 10 --warmup 3 "$@" \
      > gpurun_out/${tag}_bench_10M_${N}gpu.json 2> gpurun_out/${tag}_bench_10M_${N}gpu.log; echo "bench rc=$?"
fi
python - "$tag" "$N" <<'PY'
import json, sys
tag, N = sys.argv[1], sys.argv[2]
try:
    b = json.load(open("gpurun_out/%s_bench_10M_%sgpu.json" % (tag, N)))
    print("10M N=%s value %.0f e2e %.0f ms %.3f K1 %.3f frac %.3f" % (N, b["value"], b["e2e"]["value"], b["ms_per_step"], b["roofline"]["kernel_ms_avg"], b["roofline"]["frac"]), (b.get("hnsw") or {}).get("frontier"))
    print("cpu", b.get("cpu_baseline"))
except Exception as e:
    print("unreadable:", e)
PY
