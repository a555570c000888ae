timeout 600 python -m pytest tests/test_multigpu.py -q -x -k "fused_allgather" 2>&1 | tail -5
for idle in 0 1 2 3; do
PB_AG_IDLE_ROUNDS=$idle timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/collective_gemm_bench.py 2>/dev/null | grep -v NCCL > gpurun_out/collective_gemm_2gpu_idle$idle.json; python -c "
import json; d=json.load(open('gpurun_out/collective_gemm_2gpu_idle$idle.json'))
print($idle, [(r['fused_ms'], r['speedup_vs_nccl']) for r in d['rows']])"
done
