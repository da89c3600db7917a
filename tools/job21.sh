set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r21_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r21_pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r21_bench_2gpu.json 2> gpurun_out/r21_bench_2gpu.log; echo "2gpu rc=$?"; cat gpurun_out/r21_bench_2gpu.json | cut -c1-1400; tail -3 gpurun_out/r21_bench_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r21_ref_2gpu.json 2> gpurun_out/r21_ref_2gpu.log; echo "ref 2gpu rc=$?"; cat gpurun_out/r21_ref_2gpu.json | cut -c1-300
