timeout 900 python -m pytest tests/test_multigpu.py tests/test_elastic_gpu.py -q 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 6 --warmup 3 2>/dev/null | tee gpurun_out/bench_2gpu_final.json | cut -c1-220
