#!/usr/bin/env bash
# N-GPU exact bench on the 1M corpus: one batch in flight vs two.
tag=${1:-job}; N=${2:-8}
mkdir -p gpurun_out
for s in 1 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$s bench.py --gpus $N --steps 20 --warmup 5 --streams $s --no-hnsw --no-cpu-baseline \
      > gpurun_out/${tag}_bench_${N}gpu_s$s.json 2> gpurun_out/${tag}_bench_${N}gpu_s$s.log; echo "N=$N streams=$s rc=$?"
done
python - "$tag" "$N" <<'PY'
import json, sys, glob
tag, N = sys.argv[1], sys.argv[2]
for f in sorted(glob.glob("gpurun_out/%s_bench_%sgpu_s*.json" % (tag, N))):
    try:
        b = json.load(open(f))
        print(f.split("/")[-1], "value %.0f e2e %.0f ms %.3f K1 %.3f" % (b["value"], b["e2e"]["value"], b["ms_per_step"], b["roofline"]["kernel_ms_avg"]), b["config"].get("rescored_per_query_per_shard_device_steps"), b["config"].get("unresolved_queries"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
