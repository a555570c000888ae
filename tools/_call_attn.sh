for ps in 0 2; do
echo "PSTAGES=$ps"
PB_ATTN_BWD_PSTAGES=$ps timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "flash_attention or rope_attention or qkv_gemm_rope" 2>&1 | tail -6 | cut -c1-250
PB_ATTN_BWD_PSTAGES=$ps timeout 200 python tools/attn_bench.py 2>/dev/null | tee gpurun_out/attn_bench_v7_ps$ps.json | cut -c1-330
PB_ATTN_BWD_PSTAGES=$ps timeout 100 python tools/attn_trace.py > gpurun_out/attn_trace_v7_ps$ps.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/attn_trace_v7_ps$ps.json')); print('cycles_per_tile', d['cycles_per_tile']); print([ (t['wait_for_S'], t['tmem_load'], t['math'], t['wait_p_buffer'], t['math_store']) for t in d['per_tile'][2:8]])"
done
