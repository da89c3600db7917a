set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r6_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r6_pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r6_bench_2gpu.json 2> gpurun_out/r6_bench_2gpu.log; echo "2gpu rc=$?"; cat gpurun_out/r6_bench_2gpu.json; grep "\[bench\]" gpurun_out/r6_bench_2gpu.log | tail
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r6_ref.json 2> gpurun_out/r6_ref.log; echo "ref rc=$?"; cat gpurun_out/r6_ref.json
