set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r13_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r13_smoke.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded or shard" > gpurun_out/r13_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r13_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --workload hnsw --gpus 2 --steps 5 --warmup 2 > gpurun_out/r13_hnsw_2gpu.json 2> gpurun_out/r13_hnsw_2gpu.log; echo "hnsw 2gpu rc=$?"; cat gpurun_out/r13_hnsw_2gpu.json; grep "\[bench\]" gpurun_out/r13_hnsw_2gpu.log | tail -6; tail -5 gpurun_out/r13_hnsw_2gpu.log
