timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 6 --warmup 3 2>gpurun_out/bench8.err | tee gpurun_out/bench_8gpu_final.json | cut -c1-300; tail -3 gpurun_out/bench8.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 tools/collective_gemm_bench.py 2>gpurun_out/cg8.err | grep -v "^NCCL" > gpurun_out/collective_gemm_8gpu.json; tail -2 gpurun_out/cg8.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/collective_gemm_8gpu.json'))
for r in d['rows']: print({k:v for k,v in r.items() if 'ms' in k or 'speed' in k or k in ('M','N','K')})"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 4 --steps 6 --warmup 3 2>/dev/null | tee gpurun_out/bench_4gpu_final.json | cut -c1-200
